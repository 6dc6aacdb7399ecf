"""The algebra of the planned per-type training kernels (scripts/prototype_pertype_train.py) against the tape
of the training oracle for one ``TypeAwareMPNLayer`` -- CPU only.  The oracle itself is pinned to the reference under float64
autograd (tests/test_oracle_train.py, train_flagship.npz)."""
import importlib.util
import os

import numpy as np
import pytest

import oracle
import oracle.mpn_train as T
import pgmp_b200
import pgmp_b200.synthetic as synthetic
from cases import GC_CASES, gc_config_for, mpn_config_for
from pgmp_b200.Models.MessagePassingNetwork import get_mpn_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("prototype_pertype_train", os.path.join(ROOT, "scripts", "prototype_pertype_train.py"))
proto = importlib.util.module_from_spec(spec)
spec.loader.exec_module(proto)


@pytest.mark.parametrize("aggr_sub", ["node_edge_attn", "node_edge_attn_per_type"])
def test_planned_per_type_layer_matches_oracle_tape(aggr_sub):
    inp_kw, cfg_over = GC_CASES["tiny_complete"]
    data = synthetic.synth_batch(**inp_kw)
    g = oracle.gc.construct_graph(data["scoremaps"], data["tagmaps"], data["features"], gc_config_for(pgmp_b200.config, cfg_over),
                                  inp_kw["num_joints"], masks=data["masks"])
    cfg = mpn_config_for(pgmp_b200.config, "flagship_mpn_config", dict(STEPS=1, AGGR_SUB=aggr_sub))
    model = synthetic.synth_mpn_state_dict(get_mpn_model(cfg), 5)
    sd = {k: v.numpy().astype(np.float64) for k, v in model.state_dict().items()}
    src, dst = g["edge_index"]
    # joint types 0..3 of the 4-joint graph spread over the 17 message MLPs
    types = (g["joint_det"][:, 2] * 5) % 17
    N, E = len(types), len(src)
    rng = np.random.default_rng(0)
    x, e = rng.standard_normal((N, 128)), rng.standard_normal((E, 128))
    d_h, d_g = rng.standard_normal((N, 64)), rng.standard_normal((E, 64))
    # oracle tape
    tm = T.TrainModel(sd)
    xv, ev = T.Var(x), T.Var(e)
    h_o, g_o = tm.type_aware_layer("mpn_node_cls", xv, ev, src, dst, types, "add", aggr_sub, 17)
    loss = T.add_scalars([T.weighted_sum(h_o, d_h), T.weighted_sum(g_o, d_g)])
    T.backward(loss)
    # planned algorithm
    P = {k[len("mpn_node_cls."):]: v for k, v in sd.items() if k.startswith("mpn_node_cls.")}
    h_p, g_p, dx, de, G = proto.layer_forward_backward(P, x, e, src, dst, types, 17, aggr_sub == "node_edge_attn_per_type", d_h, d_g)

    def close(a, b, what):
        assert np.abs(a - b).max() <= 1e-9 * max(1.0, np.abs(b).max()), what
    close(h_p, h_o.value, "h'")
    close(g_p, g_o.value, "g'")
    close(dx, xv.grad, "dx")
    close(de, ev.grad, "de")
    for name, grad in G.items():
        want = tm.p["mpn_node_cls." + name].grad
        close(grad, want if want is not None else np.zeros_like(grad), name)
