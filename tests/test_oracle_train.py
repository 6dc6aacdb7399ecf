"""Round-2 groundwork: the training-mode oracle (forward with BatchNorm batch statistics + reverse-mode gradients)
against the reference run under torch autograd (tests/golden/train_*.npz)."""
import os

import numpy as np
import pytest
import torch

import pgmp_b200
import pgmp_b200.synthetic as synthetic
import oracle
import oracle.mpn_train as T
from golden.cases import GC_CASES, TRAIN_CASES, gc_config_for, mpn_config_for, sample_indices, train_loss_weights
from pgmp_b200.Models.MessagePassingNetwork import get_mpn_model

HERE = os.path.dirname(os.path.abspath(__file__))


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


@pytest.mark.parametrize("name", list(TRAIN_CASES))
def test_training_oracle_matches_reference_autograd(name):
    gc_name, maker, over, seed = TRAIN_CASES[name]
    gold = np.load(os.path.join(HERE, "golden", f"train_{name}.npz"))
    inp_kw, cfg_over = GC_CASES[gc_name]
    data = synthetic.synth_batch(**inp_kw)
    g = oracle.gc.construct_graph(data["scoremaps"], data["tagmaps"], data["features"], gc_config_for(pgmp_b200.config, cfg_over),
                                  inp_kw["num_joints"], masks=data["masks"])
    cfg = mpn_config_for(pgmp_b200.config, maker, over)
    model = synthetic.synth_mpn_state_dict(get_mpn_model(cfg), seed)
    sd = {k: v.numpy().astype(np.float64) for k, v in model.state_dict().items()}
    n_edge, n_node = int(gold["n_edge"]), int(gold["n_node"])
    shapes = [gold[f"edge_{i}"].shape for i in range(n_edge)] + [gold[f"node_{i}"].shape for i in range(n_node - 1)] + \
             [gold[f"class_{i}"].shape for i in range(n_node - 1)]
    coeffs = train_loss_weights([tuple(s) for s in shapes], seed)
    pe, pn, pc, loss, gx, grads, stats = T.loss_and_gradients(sd, cfg, g["x"], g["edge_attr"], g["edge_index"],
                                                              g["joint_det"][:, 2], coeffs)
    for i, a in enumerate(pe):
        assert rel(a, gold[f"edge_{i}"]) < 1e-6, f"edge_{i}"
    for i, a in enumerate(pn):
        assert rel(a, gold[f"node_{i}"]) < 1e-6 and rel(pc[i], gold[f"class_{i}"]) < 1e-6, f"node/class_{i}"
    assert abs(loss - float(gold["loss"])) <= 1e-9 * max(1.0, abs(float(gold["loss"])))
    assert rel(gx, gold["grad_x"]) < 1e-6, "grad_x"
    assert float(gold["fp32_grad_x_l2rel"]) < 2e-2      # the float32 reference run's own deviation, for the record
    checked = 0
    for pname, gr in grads.items():
        key = "gnorm/" + pname
        if key not in gold.files:
            continue
        want_norm = float(gold[key])
        got = gr.ravel()
        assert abs(np.linalg.norm(got) - want_norm) <= 1e-8 * max(want_norm, 1.0), pname
        samp = gold["gsamp/" + pname]
        scale = max(np.abs(samp).max(), want_norm / np.sqrt(got.size), 1e-9)
        assert np.abs(got[sample_indices(got.size, pname)] - samp).max() <= 1e-8 * max(scale, 1.0), pname
        checked += 1
    assert checked >= 20
    for bname, val in stats.items():
        assert rel(val, gold["buf/" + bname]) < 1e-9, bname
