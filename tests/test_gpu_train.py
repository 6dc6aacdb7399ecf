"""GPU parity of the training step: ``pgmp_mpn_train_forward`` / ``pgmp_mpn_train_backward``
through the module's ``train()``-mode ``forward`` + ``loss.backward()``, against the float64 training oracle
(``oracle/mpn_train.py``) and against the reference's own float64-autograd run (tests/golden/train_agnostic_max.npz).

Tolerances.  Logits and BatchNorm statistics: TIGHT everywhere.  Gradients: the test loss has random-sign
coefficients, so a gradient is a sum of E (or N) random-sign terms of norm ~ sqrt(E) and ONE ReLU / max decision that
flips between float32 and float64 (a pre-activation within 1e-7 of zero) moves it by ~ 1 / sqrt(32 E) ~ 1e-3 -- the
float32 reference run itself deviates from its float64 run by ``fp32_grad_x_l2rel`` = 3e-3 ... 5e-3 (stored in the
fixture) for the same reason.  Hence: small graphs (no flip expected) are held to TIGHT on every gradient, which pins
the algebra of every stage; the larger graphs (multi-tile, multi-split paths) to LOOSE, measured against the
reference's own float32 deviation."""
import os

import numpy as np
import pytest
import torch

import oracle
import oracle.mpn_train as T
import pgmp_b200
import pgmp_b200.synthetic as synthetic
from cases import GC_CASES, TRAIN_CASES, gc_config_for, mpn_config_for, sample_indices, train_loss_weights
from helpers import rel_err
from pgmp_b200.Models.MessagePassingNetwork import get_mpn_model

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
HERE = os.path.dirname(os.path.abspath(__file__))
TIGHT = 2e-5           # relative (linf against the tensor's scale, and l2), fp32 kernels vs the float64 oracle
LOOSE = 5e-2           # gradients on graphs large enough for ReLU / max decisions to flip in float32 (l2; linf 3x)
TINY = dict(NUM_JOINTS=4, EDGE_INPUT_DIM=6)

VARIANTS = {   # name -> (graph, MPN config overrides, weight seed, gradient tolerance)
    # small graphs (N = 40, E = 760): every stage's algebra at float32 round-off
    "tiny_max_skip_aux": ("tiny_complete", dict(STEPS=3, AUX_LOSS_STEPS=1, **TINY), 46, TIGHT),
    "tiny_add_noskip_update_mlp": ("tiny_complete", dict(AGGR="add", SKIP=False, USE_NODE_UPDATE_MLP=True, STEPS=2, **TINY), 47, TIGHT),
    "tiny_mean_all_steps": ("tiny_complete", dict(AGGR="mean", STEPS=2, AUX_LOSS_STEPS=5, **TINY), 44, TIGHT),
    "tiny_max_update_mlp": ("tiny_complete", dict(STEPS=4, USE_NODE_UPDATE_MLP=True, **TINY), 49, TIGHT),
    # the fixture case: class_agnostic_end2end shape (max aggregation, skip), one auxiliary step; N = 340, E = 20 048
    "agnostic_max": ("knn_small", dict(STEPS=3, AUX_LOSS_STEPS=1), 41, LOOSE),
    "add_noskip_update_mlp": ("knn_small", dict(AGGR="add", SKIP=False, USE_NODE_UPDATE_MLP=True, STEPS=2), 43, LOOSE),
    "max_fully_update_mlp": ("fully_small", dict(STEPS=2, USE_NODE_UPDATE_MLP=True), 45, LOOSE),     # E = 57 460
    # TypeAwareMPNLayer (layers.py:157-274; the hybrid_* configs): per-type message matrices, per-(target, type) bins
    "tiny_pt_attn": ("tiny_complete", dict(STEPS=3, AUX_LOSS_STEPS=1, **TINY), 51, TIGHT, "flagship_mpn_config"),
    "tiny_pt_attn_per_type_noskip": ("tiny_complete", dict(AGGR_SUB="node_edge_attn_per_type", SKIP=False, STEPS=2, **TINY), 52, TIGHT,
                                     "flagship_mpn_config"),
    "tiny_pt_vanilla_max": ("tiny_complete", dict(AGGR_SUB="None", AGGR="max", STEPS=2, **TINY), 53, TIGHT, "flagship_mpn_config"),
    "tiny_pt_vanilla_mean": ("tiny_complete", dict(AGGR_SUB="None", AGGR="mean", STEPS=2, AUX_LOSS_STEPS=1, **TINY), 54, TIGHT,
                             "flagship_mpn_config"),
    "pt_flagship": ("knn_small", dict(STEPS=2), 42, LOOSE, "flagship_mpn_config"),                     # the fixture case
    "pt_attn_per_type_fully": ("fully_small", dict(AGGR_SUB="node_edge_attn_per_type", STEPS=2, AUX_LOSS_STEPS=1), 55, LOOSE,
                               "flagship_mpn_config"),
}


def graph_for(gc_name):
    inp_kw, cfg_over = GC_CASES[gc_name]
    data = synthetic.synth_batch(**inp_kw)
    return oracle.gc.construct_graph(data["scoremaps"], data["tagmaps"], data["features"], gc_config_for(pgmp_b200.config, cfg_over),
                                     inp_kw["num_joints"], masks=data["masks"])


def run_cuda(cfg, seed, g):
    model = synthetic.synth_mpn_state_dict(get_mpn_model(cfg), seed)
    sd0 = {k: v.numpy().astype(np.float64) for k, v in model.state_dict().items()}
    model = model.to(DEV).train()
    x = torch.from_numpy(g["x"]).to(DEV).requires_grad_(True)
    pe, pn, pc, _ = model(x, torch.from_numpy(g["edge_attr"]).to(DEV), torch.from_numpy(g["edge_index"]).to(DEV),
                          node_types=torch.from_numpy(g["joint_det"][:, 2]).to(DEV))
    assert len(pn) == len(pe) + 1 and len(pc) == len(pe) + 1
    preds = list(pe) + list(pn[:-1]) + list(pc[:-1])
    coeffs = train_loss_weights([tuple(p.shape) for p in preds], seed)
    loss = sum((p * torch.from_numpy(c).to(DEV)).sum() for p, c in zip(preds, coeffs))
    loss.backward()
    torch.cuda.synchronize()
    return model, sd0, x, pe, pn, pc, coeffs, loss


def check(what, a, b, tol=TIGHT, linf_factor=1.0):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    assert a.shape == np.asarray(b).shape, (what, a.shape, np.asarray(b).shape)
    linf, l2 = rel_err(a, b)
    if os.environ.get("PGMP_TRAIN_REPORT"):
        print(f"    {what}: rel linf {linf:.3e} l2 {l2:.3e}")
        return l2
    assert linf <= tol * linf_factor and l2 <= tol, f"{what}: rel linf {linf:.3e} l2 {l2:.3e} > {tol:.1e}"
    return l2


def check_grad(what, a, b, tol):
    return check(what, a, b, tol, 1.0 if tol == TIGHT else 3.0)


@pytest.mark.parametrize("name", list(VARIANTS))
def test_training_step_matches_oracle(name):
    gc_name, over, seed, gtol = VARIANTS[name][:4]
    g = graph_for(gc_name)
    cfg = mpn_config_for(pgmp_b200.config, VARIANTS[name][4] if len(VARIANTS[name]) > 4 else "agnostic_mpn_config", over)
    model, sd0, x, pe, pn, pc, coeffs, loss = run_cuda(cfg, seed, g)
    ope, opn, opc, oloss, ogx, ograds, ostats = T.loss_and_gradients(sd0, cfg, g["x"], g["edge_attr"], g["edge_index"],
                                                                      g["joint_det"][:, 2], coeffs)
    worst = 0.0
    # logits: TIGHT; the per-type layer on the larger graphs (attention bins, K = 17 x 64 update product) 5e-5
    ltol = 5e-5 if len(VARIANTS[name]) > 4 and gtol == LOOSE else TIGHT
    for i in range(len(ope)):
        worst = max(worst, check(f"edge_{i}", pe[i], ope[i], ltol), check(f"node_{i}", pn[i], opn[i], ltol),
                    check(f"class_{i}", pc[i], opc[i], ltol))
    assert torch.equal(pn[-1], pn[-2]) and torch.equal(pc[-1], pc[-2])          # NodeClassificationMPNSimple.py:93-94
    worst = max(worst, check_grad("grad_x", x.grad, ogx, gtol))
    params = dict(model.named_parameters())
    assert set(params) == set(ograds)
    for pname, want in ograds.items():
        got = params[pname].grad
        assert got is not None, pname
        if np.abs(want).max() < 1e-9:
            # exactly zero in exact arithmetic (e.g. attn_net bias: a softmax ignores a shift of its bin's logits); in
            # float32 a cancelling sum over the E edges leaves round-off
            assert float(got.abs().max()) < 2e-6 * np.sqrt(max(g["edge_index"].shape[1], 100)), pname
            continue
        worst = max(worst, check_grad("grad " + pname, got, want, gtol))
    for bname, want in ostats.items():
        check("buffer " + bname, model.state_dict()[bname], want, 1e-5)
    for mod in model.modules():
        if isinstance(mod, torch.nn.BatchNorm1d):
            assert int(mod.num_batches_tracked) == 1
    print(f"{name}: worst relative deviation from the float64 oracle {worst:.2e}")


@pytest.mark.parametrize("case", ["agnostic_max", "flagship"])
def test_training_step_matches_reference_fixture(case):
    """Against the UNMODIFIED reference under float64 autograd (tests/golden/make_golden_train.py)."""
    gc_name, maker, over, seed = TRAIN_CASES[case]
    gold = np.load(os.path.join(HERE, "golden", "train_%s.npz" % case))
    g = graph_for(gc_name)
    cfg = mpn_config_for(pgmp_b200.config, maker, over)
    model, sd0, x, pe, pn, pc, coeffs, loss = run_cuda(cfg, seed, g)
    ltol = TIGHT if case == "agnostic_max" else 5e-5       # per-type layer: see test_training_step_matches_oracle
    for i, a in enumerate(pe):
        check(f"edge_{i}", a, gold[f"edge_{i}"], ltol)
    for i in range(len(pn)):
        check(f"node_{i}", pn[i], gold[f"node_{i}"], ltol)
        check(f"class_{i}", pc[i], gold[f"class_{i}"], ltol)
    assert abs(float(loss) - float(gold["loss"])) <= 1e-4 * max(1.0, abs(float(gold["loss"])))
    check_grad("grad_x", x.grad, gold["grad_x"], LOOSE)
    # the yardstick: the reference's own float32 run against its float64 run (same ReLU / max decision flips)
    _, l2 = rel_err(x.grad.cpu().numpy(), gold["grad_x"])
    assert l2 < 3 * float(gold["fp32_grad_x_l2rel"]), (l2, float(gold["fp32_grad_x_l2rel"]))
    checked = 0
    for pname, p in model.named_parameters():
        got = p.grad.cpu().numpy().ravel().astype(np.float64)
        want_norm = float(gold["gnorm/" + pname])
        if want_norm < 1e-9:     # zero in exact arithmetic (attn_net bias, unused type matrices): float32 round-off of a cancelling sum
            assert np.abs(got).max() < 2e-6 * np.sqrt(g["edge_index"].shape[1]), pname
            checked += 1
            continue
        assert abs(np.linalg.norm(got) - want_norm) <= LOOSE * max(want_norm, 1e-6), pname
        samp = gold["gsamp/" + pname]
        scale = max(np.abs(samp).max(), want_norm / np.sqrt(got.size), 1e-9)
        assert np.abs(got[sample_indices(got.size, pname)] - samp).max() <= 5 * LOOSE * scale, pname
        checked += 1
    assert checked >= 20
    for bname in (k[4:] for k in gold.files if k.startswith("buf/")):
        check("buffer " + bname, model.state_dict()[bname], gold["buf/" + bname], 1e-5)


def test_training_step_is_reproducible_and_eval_still_works():
    """No floating-point atomics in the reverse pass: two runs give identical bits; the updated running statistics
    are what the inference path then folds."""
    gc_name, over, seed, _ = VARIANTS["agnostic_max"]
    g = graph_for(gc_name)
    cfg = mpn_config_for(pgmp_b200.config, "agnostic_mpn_config", over)
    a = run_cuda(cfg, seed, g)
    b = run_cuda(cfg, seed, g)
    assert torch.equal(a[2].grad, b[2].grad)
    for (n1, p1), (_, p2) in zip(a[0].named_parameters(), b[0].named_parameters()):
        assert torch.equal(p1.grad, p2.grad), n1
    model = a[0].eval()
    with torch.no_grad():
        pe, pn, pc, _ = model(a[2].detach(), torch.from_numpy(g["edge_attr"]).to(DEV), torch.from_numpy(g["edge_index"]).to(DEV),
                              node_types=torch.from_numpy(g["joint_det"][:, 2]).to(DEV))
    sd = {k: v.cpu().numpy() for k, v in model.state_dict().items()}
    ope, opn, opc = oracle.mpn.node_classification_mpn_forward(sd, cfg, g["x"], g["edge_attr"], g["edge_index"], g["joint_det"][:, 2])
    check("eval edge", pe[-1], ope[-1], 1e-4)
    check("eval node", pn[-1], opn[-1], 1e-4)


def test_training_hierarch_update_raises():
    cfg = mpn_config_for(pgmp_b200.config, "flagship_mpn_config", dict(STEPS=2, UPDATE_TYPE="hierarch_mlp"))
    model = get_mpn_model(cfg).to(DEV).train()
    g = graph_for("knn_small")
    with pytest.raises(NotImplementedError):
        model(torch.from_numpy(g["x"]).to(DEV), torch.from_numpy(g["edge_attr"]).to(DEV), torch.from_numpy(g["edge_index"]).to(DEV),
              node_types=torch.from_numpy(g["joint_det"][:, 2]).to(DEV))


@pytest.mark.parametrize("maker", ["agnostic_mpn_config", "flagship_mpn_config"])
def test_training_full_size_properties(maker):
    """BASELINE configs[4] size (8 images of 256 x 256, ~231 k edges, 10 steps), where the oracle is too slow:
    size-independent properties of the reverse pass.  It is linear in dL/d(logits) -- grads(d1 + d2) = grads(d1) + grads(d2),
    grads(0) = 0 -- and running it twice on the same forward gives identical bits."""
    from pgmp_b200.graph_constructor import get_graph_constructor
    J, K, B = 17, 30, 8
    data = synthetic.synth_batch(B, J, 256, K, persons=8)
    t = {k: torch.from_numpy(v).to(DEV) for k, v in data.items()}
    ret = get_graph_constructor(pgmp_b200.config.bench_gc_config(k=K, graph_type="knn"), scoremaps=t["scoremaps"], tagmaps=t["tagmaps"],
                                features=t["features"], joints_gt=None, factor_list=None, masks=None, device=DEV, testing=False,
                                heatmaps=None, num_joints=J).construct_graph()
    x, edge_attr, edge_index, joint_det = ret[0].clone().requires_grad_(True), ret[1], ret[2], ret[7]
    assert x.shape[0] == B * J * K and edge_index.shape[1] > 200_000
    model = synthetic.synth_mpn_state_dict(get_mpn_model(getattr(pgmp_b200.config, maker)(J, AUX_LOSS_STEPS=1)), 7).to(DEV).train()
    pe, pn, pc, _ = model(x, edge_attr, edge_index, node_types=joint_det[:, 2])
    outs = [pe[0], pe[1], pn[0], pn[1], pc[0], pc[1]]
    assert all(bool(torch.isfinite(o).all()) for o in outs)
    params = [x] + list(model.parameters())
    gen = torch.Generator(device=DEV).manual_seed(3)
    d1 = [torch.randn(o.shape, device=DEV, generator=gen) for o in outs]
    d2 = [torch.randn(o.shape, device=DEV, generator=gen) for o in outs]

    def grads(d):
        return torch.autograd.grad(outs, params, grad_outputs=d, retain_graph=True)

    g1, g2, g12, g1b = grads(d1), grads(d2), grads([a + b for a, b in zip(d1, d2)]), grads(d1)
    g0 = grads([torch.zeros_like(o) for o in outs])
    names = ["x"] + [n for n, _ in model.named_parameters()]
    for a, b, c, a2, z, p, pname in zip(g1, g2, g12, g1b, g0, params, names):
        assert torch.equal(a, a2)                                   # reproducible
        assert float(z.abs().max()) == 0.0                          # zero in, zero out
        if pname.endswith("attn_net.0.bias"):                       # zero in exact arithmetic (softmax shift invariance): pure round-off
            assert float(c.abs().max()) < 2e-6 * np.sqrt(edge_index.shape[1])
            continue
        scale = float(c.abs().max()) + 1e-30
        assert float((a + b - c).abs().max()) <= 2e-4 * scale, (tuple(p.shape), float((a + b - c).abs().max()) / scale)
    # the activations are gone once another forward has taken the pooled workspace
    model(x, edge_attr, edge_index, node_types=joint_det[:, 2])
    with pytest.raises(RuntimeError):
        grads(d1)


def test_gather_gradient_reaches_the_feature_maps():
    """End-to-end training (train.py:232): dL/dx flows back into the feature maps through pgmp_gc_gather_backward.
    Three joint types share their heatmap here, so their candidates share pixels and the kernel has to sum duplicates;
    the result is bit-identical to the sequential numpy scatter-add (same order of additions)."""
    from pgmp_b200.graph_constructor import get_graph_constructor
    J, K = 6, 8
    data = synthetic.synth_batch(3, J, 96, K, channels=40, persons=2, width=128)
    data["scoremaps"][:, 1] = data["scoremaps"][:, 0]
    data["scoremaps"][:, 2] = data["scoremaps"][:, 0]
    feat = torch.from_numpy(data["features"]).to(DEV).requires_grad_(True)
    ret = get_graph_constructor(pgmp_b200.config.bench_gc_config(k=K, graph_type="knn"), scoremaps=torch.from_numpy(data["scoremaps"]).to(DEV),
                                tagmaps=torch.from_numpy(data["tagmaps"]).to(DEV), features=feat, joints_gt=None, factor_list=None,
                                masks=None, device=DEV, testing=False, heatmaps=None, num_joints=J).construct_graph()
    x, joint_det, batch_index = ret[0], ret[7].cpu().numpy(), ret[12].cpu().numpy()
    assert x.requires_grad
    pix = {(b, y, xx) for b, (xx, y, _) in zip(batch_index, joint_det)}
    assert len(pix) < len(joint_det)                                   # duplicates are present
    coeff = np.random.default_rng(5).standard_normal(tuple(x.shape)).astype(np.float32)
    (x * torch.from_numpy(coeff).to(DEV)).sum().backward()
    want = np.zeros(data["features"].shape, np.float32)
    for n in range(len(joint_det)):                                    # sequential, node order
        want[batch_index[n], :, joint_det[n, 1], joint_det[n, 0]] += coeff[n]
    assert np.array_equal(feat.grad.cpu().numpy(), want)
    # channels-last feature maps (the strides are honoured)
    feat2 = torch.from_numpy(data["features"]).to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    ret2 = get_graph_constructor(pgmp_b200.config.bench_gc_config(k=K, graph_type="knn"), scoremaps=torch.from_numpy(data["scoremaps"]).to(DEV),
                                 tagmaps=torch.from_numpy(data["tagmaps"]).to(DEV), features=feat2, joints_gt=None, factor_list=None,
                                 masks=None, device=DEV, testing=False, heatmaps=None, num_joints=J).construct_graph()
    (ret2[0] * torch.from_numpy(coeff).to(DEV)).sum().backward()
    assert np.array_equal(feat2.grad.cpu().numpy(), want)


@pytest.mark.parametrize("name", ["agnostic_max", "max_fully_update_mlp", "pt_flagship"])
def test_training_forward_products_on_tcgen05(name, monkeypatch):
    """PGMP_TRAIN_TC=1: the E-level forward products run on the 5th-generation tensor cores (mpn_train_tc.cu, bf16 hi / lo
    operand pairs, fp32 accumulation in tensor memory).  Logits within 1e-4 of the float64 oracle (the 3xTF32 default:
    2e-5), gradients within the same bound as the default mode; the kernel must actually have run."""
    import pgmp_b200._native as nv
    monkeypatch.setenv("PGMP_TRAIN_TC", "1")
    gc_name, over, seed, gtol = VARIANTS[name][:4]
    g = graph_for(gc_name)
    cfg = mpn_config_for(pgmp_b200.config, VARIANTS[name][4] if len(VARIANTS[name]) > 4 else "agnostic_mpn_config", over)
    nv.profile(True)
    model, sd0, x, pe, pn, pc, coeffs, loss = run_cuda(cfg, seed, g)
    prof = nv.profile_collect()
    nv.profile(False)
    assert any(k.startswith("lin_fwd_tc_kernel") for k in prof), sorted(prof)
    ope, opn, opc, oloss, ogx, ograds, ostats = T.loss_and_gradients(sd0, cfg, g["x"], g["edge_attr"], g["edge_index"],
                                                                      g["joint_det"][:, 2], coeffs)
    for i in range(len(ope)):
        check(f"edge_{i}", pe[i], ope[i], 1e-4)
        check(f"node_{i}", pn[i], opn[i], 1e-4)
        check(f"class_{i}", pc[i], opc[i], 1e-4)
    check_grad("grad_x", x.grad, ogx, gtol)
    params = dict(model.named_parameters())
    for pname, want in ograds.items():
        if np.abs(want).max() >= 1e-9:
            check_grad("grad " + pname, params[pname].grad, want, gtol)
