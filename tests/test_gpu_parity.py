"""GPU parity tests (B200): the CUDA path, called through the reference-shaped host API and the
C ABI, against the numpy oracle on the same seeded inputs and against the reference's own outputs
(tests/golden).  Integer / index / gathered quantities are bit-exact; logits within tolerance."""
import copy

import numpy as np
import pytest
import torch

import oracle
import pgmp_b200
import pgmp_b200.synthetic as synthetic
from cases import FULL_CASES, full_edge_sample
from helpers import (LOGIT_TOL, FP32_TOL, FP32_TOL_FULL, GC_CASES, GC_KEYS, MPN_CASES, assert_close, assert_matches_golden,
                     gc_config_for, gc_inputs, golden, mpn_config_for)
from pgmp_b200.graph_constructor import get_graph_constructor
from pgmp_b200.Models.MessagePassingNetwork import get_mpn_model

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SLOT = dict(x=0, edge_attr=1, edge_index=2, joint_det=7, joint_scores=11, batch_index=12, joint_tags=14)


def run_gc(name, **extra):
    data, cfg, nj = gc_inputs(name)
    cfg = copy.copy(cfg)
    for k, v in extra.items():
        setattr(cfg, k, v)
    t = {k: torch.from_numpy(v).to(DEV) for k, v in data.items()}
    gc = get_graph_constructor(cfg, scoremaps=t["scoremaps"], tagmaps=t["tagmaps"], features=t["features"],
                               joints_gt=None, factor_list=None, masks=t["masks"] if cfg.MASK_CROWDS else None,
                               device=DEV, testing=True, heatmaps=None, num_joints=nj)
    ret = gc.construct_graph()
    assert len(ret) == 15 and all(ret[i] is None for i in (3, 4, 5, 6, 8, 9, 10, 13))
    return ret, gc


@pytest.mark.parametrize("name", list(GC_CASES))
def test_graph_constructor_bit_exact(name):
    ret, gc = run_gc(name)
    data, cfg, nj = gc_inputs(name)
    want = oracle.gc.construct_graph(data["scoremaps"], data["tagmaps"], data["features"], cfg, nj, masks=data["masks"])
    gold = golden("gc_" + name)
    for k in GC_KEYS:
        got = ret[SLOT[k]].cpu().numpy()
        assert got.dtype == want[k].dtype and got.shape == want[k].shape, (k, got.shape, want[k].shape)
        assert np.array_equal(got, want[k]), f"{k}: {np.sum(got != want[k])} mismatches vs oracle"
        assert_matches_golden(gold, k, got, exact=True)
    assert np.array_equal(gc.num_nodes_per_image.numpy(), want["num_nodes"])
    assert np.array_equal(gc.num_edges_per_image.numpy(), want["num_edges"])


def test_graph_constructor_channels_last_and_5d_tags():
    data, cfg, nj = gc_inputs("knn_small")
    sm = torch.from_numpy(data["scoremaps"]).to(DEV)
    feat = torch.from_numpy(data["features"]).to(DEV).contiguous(memory_format=torch.channels_last)
    tags5 = torch.from_numpy(np.stack([data["tagmaps"], -data["tagmaps"]], -1)).to(DEV)
    ret = get_graph_constructor(cfg, scoremaps=sm, tagmaps=tags5, features=feat, joints_gt=None, factor_list=None,
                                masks=None, device=DEV, testing=True, heatmaps=None, num_joints=nj).construct_graph()
    want = oracle.gc.construct_graph(data["scoremaps"], tags5.cpu().numpy(), data["features"], cfg, nj)
    assert np.array_equal(ret[0].cpu().numpy(), want["x"])
    assert np.array_equal(ret[14].cpu().numpy(), want["joint_tags"]) and ret[14].shape[1] == 2


def test_graph_constructor_reads_pinned_host_maps_in_place():
    """Pinned host feature / tag maps are gathered in place over PCIe; results are bit-identical to device-resident
    inputs, for NCHW and channels-last strides; pageable host tensors are moved to the device as the reference does."""
    data, gcfg, nj = gc_inputs("knn_small")
    want, _ = run_gc("knn_small")
    sm = torch.from_numpy(data["scoremaps"]).to(DEV)
    for fmt in (torch.contiguous_format, torch.channels_last):
        feat_h = torch.from_numpy(data["features"]).contiguous(memory_format=fmt).pin_memory()
        tags_h = torch.from_numpy(data["tagmaps"]).pin_memory()
        masks = torch.from_numpy(data["masks"]).to(DEV) if gcfg.MASK_CROWDS else None
        gc = get_graph_constructor(gcfg, scoremaps=sm, tagmaps=tags_h, features=feat_h, joints_gt=None, factor_list=None,
                                   masks=masks, device=DEV, testing=True, heatmaps=None, num_joints=nj)
        assert gc.features.device.type == "cpu" and gc.tagmaps.device.type == "cpu"
        got = gc.construct_graph()
        torch.cuda.synchronize()
        for i in (0, 1, 2, 7, 11, 12, 14):
            assert got[i].device.type == "cuda" and torch.equal(got[i], want[i]), i
    gc = get_graph_constructor(gcfg, scoremaps=sm, tagmaps=torch.from_numpy(data["tagmaps"]),
                               features=torch.from_numpy(data["features"]), joints_gt=None, factor_list=None, masks=masks,
                               device=DEV, testing=True, heatmaps=None, num_joints=nj)
    assert gc.features.device.type == "cuda"
    got = gc.construct_graph()
    assert torch.equal(got[0], want[0]) and torch.equal(got[14], want[14])


def test_graph_constructor_capacity_errors_are_loud():
    with pytest.raises(RuntimeError, match="B200_MAX_NODES"):
        run_gc("knn_small", B200_MAX_NODES=64)
    with pytest.raises(RuntimeError, match="B200_MAX_DET_PER_TYPE"):
        run_gc("thr_extras", B200_MAX_DET_PER_TYPE=8)


def _gc_vs_oracle(x, cfg, nj, masks=None, feat_c=2):
    """construct_graph() of the CUDA path on scoremaps ``x`` [B,J,H,W] against the oracle, every output bit-exact."""
    rng = np.random.default_rng(int(x.size) % 9973)
    feat = rng.standard_normal((x.shape[0], feat_c) + x.shape[2:]).astype(np.float32)
    tags = rng.standard_normal(x.shape).astype(np.float32)
    t = lambda a: torch.from_numpy(a).to(DEV)
    ret = get_graph_constructor(cfg, scoremaps=t(x), tagmaps=t(tags), features=t(feat), joints_gt=None, factor_list=None,
                                masks=t(masks) if masks is not None else None, device=DEV, testing=True, heatmaps=None,
                                num_joints=nj).construct_graph()
    want = oracle.gc.construct_graph(x, tags, feat, cfg, nj, masks=masks)
    for k in GC_KEYS:
        got = ret[SLOT[k]].cpu().numpy()
        assert got.shape == want[k].shape, (k, got.shape, want[k].shape)
        assert np.array_equal(got, want[k]), f"{k}: {np.sum(got != want[k])} mismatches vs oracle"
    return ret


@pytest.mark.parametrize("H,W,pool,strip", [(64, 512, 5, 0), (200, 640, 5, 32), (77, 333, 3, 16), (40, 1100, 5, 0),
                                            (96, 2052, 7, 40), (33, 50, 9, 8), (130, 130, 1, 64), (512, 512, 5, 512)])
def test_nms_shapes_strips_and_tiles(H, W, pool, strip, monkeypatch):
    """The shared-memory row-ring NMS over map shapes that exercise every loader: one contiguous bulk copy per stage
    (W % 4 == 0, one tile), per-row bulk copies (W > 1024: several column tiles with halos), thread copies (W % 4 != 0),
    all pool kernels, strips that end inside a stage, and the running cut with several strips per map."""
    if strip:
        monkeypatch.setenv("PGMP_NMS_STRIP_ROWS", str(strip))
    rng = np.random.default_rng(H * 7 + W)
    J = 3
    x = rng.uniform(0, 0.02, (2, J, H, W)).astype(np.float32)
    for b in range(2):
        for j in range(J):
            for _ in range(40):       # distinct peaks incl. the borders and corners
                y, xx = int(rng.integers(0, H)), int(rng.integers(0, W))
                x[b, j, y, xx] = rng.uniform(0.1, 1.0)
    x[0, 0, 0, 0] = x[0, 0, H - 1, W - 1] = x[0, 0, 0, W - 1] = x[0, 0, H - 1, 0] = 0.97
    x[1, 1, H // 2, :7] = 0.5         # a plateau: every pixel of it is a maximum
    cfg = pgmp_b200.config.bench_gc_config(k=12, POOL_KERNEL_SIZE=pool, DETECT_THRESHOLD=0.9, graph_type="knn")
    if pool == 1:                     # every positive pixel is a maximum
        cfg = pgmp_b200.config.bench_gc_config(k=12, POOL_KERNEL_SIZE=1, DETECT_THRESHOLD=0.99, graph_type="knn")
    _gc_vs_oracle(x, cfg, J)


def test_nms_crowd_mask_with_fractional_and_large_values():
    """score = x * mask (CG.py:1163-1165) with mask values other than 0 / 1: the running cut compares the masked score."""
    rng = np.random.default_rng(11)
    x = rng.uniform(0, 0.3, (1, 4, 96, 128)).astype(np.float32)
    masks = rng.choice(np.array([0.0, 0.5, 1.0, 3.0], np.float32), size=(1, 96, 128))
    cfg = pgmp_b200.config.bench_gc_config(k=9, POOL_KERNEL_SIZE=3, DETECT_THRESHOLD=0.6, MASK_CROWDS=True, graph_type="knn")
    _gc_vs_oracle(x, cfg, 4, masks=masks)


def test_no_threshold_path_pads_with_zero_score_pixels():
    """Fewer than 20 positive maxima for a joint (clean or crowd-masked maps): torch.topk pads the block with zero-score
    pixels and `+ 1e-10` keeps them (CG.py:1184-1195) -- no error, exactly 20 J rows."""
    rng = np.random.default_rng(3)
    J, H, W = 4, 48, 64
    x = np.zeros((2, J, H, W), np.float32)
    for j, n in enumerate((25, 4, 0, 19)):
        for _ in range(n):
            x[0, j, int(rng.integers(0, H)), int(rng.integers(0, W))] = rng.uniform(0.1, 1.0)
    x[0, 1, 0, 0] = 0.4                                   # the first pixels are maxima themselves: not padding
    x[0, 1, 0, 2] = 0.3
    x[1] = rng.uniform(0, 1, (J, H, W)).astype(np.float32)
    masks = np.ones((2, H, W), np.float32)
    masks[1, :, 5:] = 0                                   # image 1: only 5 columns survive the crowd mask
    for mc in (False, True):
        cfg = pgmp_b200.config.bench_gc_config(k=5, DETECT_THRESHOLD=2.0, MASK_CROWDS=mc, graph_type="knn")
        ret = _gc_vs_oracle(x, cfg, J, masks=masks if mc else None)
        assert ret[7].shape[0] == 2 * J * 20
        assert float(ret[11].min()) == np.float32(1e-10) or not mc


def test_default_capacities_grow_on_overflow():
    """A noisy map with 3 x 3 pooling and a low threshold has ~1 000 detections per joint: the default device capacities
    (256 per type, 1 024 nodes) grow and the detection is repeated instead of raising."""
    rng = np.random.default_rng(5)
    x = rng.uniform(0.1, 1.0, (1, 4, 96, 96)).astype(np.float32)
    cfg = pgmp_b200.config.bench_gc_config(k=5, POOL_KERNEL_SIZE=3, DETECT_THRESHOLD=0.2, graph_type="knn")
    ret = _gc_vs_oracle(x, cfg, 4)
    assert ret[7].shape[0] > 2048


def test_nms_hand_made_plateaus_borders():
    """Equal-valued positive plateaus are all maxima; borders behave like -inf padding."""
    x = np.zeros((1, 2, 16, 16), dtype=np.float32)
    x[0, 0, 0, 0] = 0.5
    x[0, 0, 5, 5] = x[0, 0, 5, 6] = 0.7
    x[0, 0, 15, 9] = 0.3
    x[0, 1, 8, 8] = 0.9
    x[0, 1, 8, 9] = 0.8          # suppressed by its neighbour
    x[0, 1, 3, 15] = 0.2
    cfg = pgmp_b200.config.bench_gc_config(k=4, POOL_KERNEL_SIZE=3, DETECT_THRESHOLD=0.6, graph_type="fully")
    t = torch.from_numpy(x).to(DEV)
    feat = torch.arange(16 * 16, dtype=torch.float32, device=DEV).reshape(1, 1, 16, 16)
    ret = get_graph_constructor(cfg, scoremaps=t, tagmaps=t, features=feat, joints_gt=None, factor_list=None,
                                masks=None, device=DEV, testing=True, heatmaps=None, num_joints=2).construct_graph()
    want = oracle.gc.construct_graph(x, x, feat.cpu().numpy(), cfg, 2)
    assert np.array_equal(ret[7].cpu().numpy(), want["joint_det"])
    assert ret[7].cpu().numpy().tolist() == [[0, 0, 0], [5, 5, 0], [6, 5, 0], [9, 15, 0], [15, 3, 1], [8, 8, 1]]
    assert np.array_equal(ret[2].cpu().numpy(), want["edge_index"])


def mpn_case(name, precision="fp32"):
    gc_name, maker, over, seed = MPN_CASES[name]
    cfg = mpn_config_for(pgmp_b200.config, maker, over)
    cfg.B200_PRECISION = precision
    data, gcfg, nj = gc_inputs(gc_name)
    g = oracle.gc.construct_graph(data["scoremaps"], data["tagmaps"], data["features"], gcfg, nj, masks=data["masks"])
    model = synthetic.synth_mpn_state_dict(get_mpn_model(cfg), seed).eval().to(DEV)
    return cfg, g, model


def run_mpn(model, g):
    with torch.no_grad():
        return model(torch.from_numpy(g["x"]).to(DEV), torch.from_numpy(g["edge_attr"]).to(DEV),
                     torch.from_numpy(g["edge_index"]).to(DEV),
                     node_types=torch.from_numpy(g["joint_det"][:, 2]).to(DEV))


@pytest.mark.parametrize("name", list(MPN_CASES))
def test_mpn_fp32_matches_oracle_and_reference(name):
    cfg, g, model = mpn_case(name)
    pe, pn, pc, tag = run_mpn(model, g)
    assert tag == [None]
    sd = {k: v.cpu().numpy() for k, v in model.state_dict().items()}
    ope, opn, opc = oracle.mpn.node_classification_mpn_forward(sd, cfg, g["x"], g["edge_attr"], g["edge_index"],
                                                               g["joint_det"][:, 2])
    gold = golden("mpn_" + name)
    assert (len(pe), len(pn), len(pc)) == (len(ope), len(opn), len(opc)) == (gold["n_edge"], gold["n_node"], gold["n_class"])
    for kind, got, want in (("edge", pe, ope), ("node", pn, opn), ("class", pc, opc)):
        for i, (a, b) in enumerate(zip(got, want)):
            a = a.cpu().numpy()
            assert_close(a, b, FP32_TOL, f"{name}:{kind}_{i} vs oracle")
            assert_close(a, gold[f"{kind}_{i}"], FP32_TOL, f"{name}:{kind}_{i} vs reference")


def test_mpn_is_deterministic_and_order_independent():
    """Same graph with the edge list shuffled gives the same logits per edge (bit-exact: slots are
    ordered by (type, target, edge id) only inside bins, so compare after un-shuffling within tolerance),
    and two runs on identical input are bit-identical."""
    cfg, g, model = mpn_case("flagship")
    a = run_mpn(model, g)
    b = run_mpn(model, g)
    for u, v in zip(a[0] + a[1] + a[2], b[0] + b[1] + b[2]):
        assert torch.equal(u, v)
    perm = np.random.default_rng(0).permutation(g["edge_index"].shape[1])
    g2 = dict(g, edge_index=g["edge_index"][:, perm], edge_attr=g["edge_attr"][perm])
    c = run_mpn(model, g2)
    assert_close(c[0][0].cpu().numpy(), a[0][0].cpu().numpy()[perm], FP32_TOL, "shuffled edges")
    assert_close(c[1][0].cpu().numpy(), a[1][0].cpu().numpy(), FP32_TOL, "shuffled nodes")


def test_mpn_tiny_graphs_and_squeeze_semantics():
    cfg = pgmp_b200.config.flagship_mpn_config(STEPS=2)
    model = synthetic.synth_mpn_state_dict(get_mpn_model(cfg), 3).eval().to(DEV)
    rng = np.random.default_rng(0)
    for n, edges in ((1, []), (2, [(0, 1)]), (2, [(0, 1), (1, 0)]), (3, [(0, 1), (1, 0), (2, 0), (0, 2)])):
        x = rng.standard_normal((n, 128)).astype(np.float32)
        ei = np.array(edges, dtype=np.int64).reshape(-1, 2).T.copy()
        ea = rng.standard_normal((ei.shape[1], 19)).astype(np.float32)
        nt = rng.integers(0, 17, size=n)
        pe, pn, pc, _ = model(torch.from_numpy(x).to(DEV), torch.from_numpy(ea).to(DEV), torch.from_numpy(ei).to(DEV),
                              node_types=torch.from_numpy(nt).to(DEV))
        sd = {k: v.cpu().numpy() for k, v in model.state_dict().items()}
        ope, opn, opc = oracle.mpn.node_classification_mpn_forward(sd, cfg, x, ea, ei, nt)
        assert pe[0].shape == ope[0].shape and pn[0].shape == opn[0].shape and pc[0].shape == opc[0].shape
        # one- or two-element tensors have no meaningful scale: absolute tolerance on O(1) logits
        np.testing.assert_allclose(pn[0].cpu().numpy(), opn[0], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(pc[0].cpu().numpy(), opc[0], rtol=1e-4, atol=1e-5)
        if ei.shape[1]:
            np.testing.assert_allclose(pe[0].cpu().numpy(), ope[0], rtol=1e-4, atol=1e-5)


def test_end_to_end_graph_constructor_into_mpn():
    """construct_graph() output feeds forward() directly, as PoseEstimationBaseline.forward does
    (PoseEstimation.py:82-93)."""
    ret, _ = run_gc("knn_small")
    cfg, g, model = mpn_case("flagship")
    with torch.no_grad():
        pe, pn, pc, _ = model(ret[0], ret[1], ret[2], node_labels=None, edge_labels=None, batch_index=ret[12],
                              node_mask=None, node_types=ret[7][:, 2].detach(), joint_tags=ret[14])
    gold = golden("mpn_flagship")
    assert_close(pe[-1].cpu().numpy(), gold["edge_0"], FP32_TOL, "edge")
    assert_close(pn[-1].cpu().numpy(), gold["node_1"], FP32_TOL, "node")
    assert_close(pc[-1].cpu().numpy(), gold["class_1"], FP32_TOL, "class")
    pe[-1] = torch.sigmoid(pe[-1])          # the caller mutates list entries (PoseEstimation.py:95-101)


# ------------------------------------------------------------------------------------------------
# tensor-core mode (tcgen05 / TMEM, bf16x3 split with fp32 accumulation)
# ------------------------------------------------------------------------------------------------
def test_pipelined_api_equals_serial_calls():
    """``GroupingPipeline`` (detection half of the next batch on a side stream, pinned host heatmaps copied there) returns
    exactly what back-to-back ``construct_graph()`` + ``forward()`` calls return, batch after batch."""
    from pgmp_b200.pipeline import GroupingPipeline
    J, K = 17, 10
    gcfg = pgmp_b200.config.bench_gc_config(k=K, graph_type="knn")
    mcfg = pgmp_b200.config.flagship_mpn_config(J, STEPS=3, B200_PRECISION="tc")
    model = synthetic.synth_mpn_state_dict(get_mpn_model(mcfg), 4).eval().to(DEV)
    datas = [synthetic.synth_batch(2, J, 128, K, persons=3, first_index=2 * i) for i in range(4)]
    serial = []
    for d in datas:
        t = {k: torch.from_numpy(v).to(DEV) for k, v in d.items()}
        ret = get_graph_constructor(gcfg, scoremaps=t["scoremaps"], tagmaps=t["tagmaps"], features=t["features"], joints_gt=None,
                                    factor_list=None, masks=None, device=DEV, testing=True, heatmaps=None, num_joints=J).construct_graph()
        with torch.no_grad():
            pe, pn, pc, _ = model(ret[0], ret[1], ret[2], node_types=ret[7][:, 2])
        serial.append((ret, pe[-1], pn[-1], pc[-1]))
    from pgmp_b200.Utils import group_persons
    serial_groups = []
    for d, (sret, spe, spn, spc) in zip(datas, serial):
        serial_groups.append(group_persons(sret[7], spn, sret[2], spe, spc, sret[12], J, node_threshold=0.5, detector_scores=sret[11]))
    pipe_g = GroupingPipeline(gcfg, model, J, DEV, group=dict(node_threshold=0.5))
    batches = [{k: torch.from_numpy(d[k]).to(DEV) for k in ("scoremaps", "tagmaps", "features")} for d in datas]
    outs = list(pipe_g.run(batches))
    assert len(outs) == len(datas)
    for (ret, preds, groups), want in zip(outs, serial_groups):
        assert len(groups) == len(want) == 2
        for g_, w_ in zip(groups, want):
            assert (g_ is None) == (w_ is None)
            if g_ is not None:
                assert np.array_equal(g_[0], w_[0]) and g_[1] == w_[1] and torch.equal(g_[2], w_[2])
    pipe = GroupingPipeline(gcfg, model, J, DEV)
    for host in (False, True):
        if host:      # pinned host inputs: heatmaps copied on the side stream, feature / tag maps gathered in place
            batches = [{k: torch.from_numpy(d[k]).pin_memory() for k in ("scoremaps", "tagmaps", "features")} for d in datas]
        else:
            batches = [{k: torch.from_numpy(d[k]).to(DEV) for k in ("scoremaps", "tagmaps", "features")} for d in datas]
        outs = list(pipe.run(batches))
        torch.cuda.synchronize()
        assert len(outs) == len(serial)
        for (ret, (pe, pn, pc)), (sret, spe, spn, spc) in zip(outs, serial):
            for i in (0, 1, 2, 7, 11, 12, 14):
                assert torch.equal(ret[i], sret[i]), i
            assert torch.equal(pe[-1], spe) and torch.equal(pn[-1], spn) and torch.equal(pc[-1], spc)


def test_malformed_edge_index_is_reported_not_a_device_fault():
    """A node id outside [0, N) in ``edge_index`` (the reference raises an IndexError): the edge is dropped, nothing is
    written out of bounds, and the module raises once the status word of that forward has arrived; ``.data`` updates are
    picked up after ``invalidate_packed_weights()``."""
    cfg, g, model = mpn_case("flagship", "tc")
    x, ea, ei, types = (torch.from_numpy(g[k]).to(DEV) for k in ("x", "edge_attr", "edge_index", "joint_det"))
    types = types[:, 2].contiguous()
    with torch.no_grad():
        good = model(x, ea, ei, node_types=types)[0][-1].clone()
    bad = ei.clone()
    bad[0, 5] = x.shape[0] + 7
    bad[1, 9] = -3
    with torch.no_grad():
        model(x, ea, bad, node_types=types)
    torch.cuda.synchronize()
    with pytest.raises(IndexError):
        model.check_status()
    with torch.no_grad():
        again = model(x, ea, ei, node_types=types)[0][-1]
    assert torch.equal(again, good)
    p0 = next(model.parameters())
    p0.data.mul_(1.5)
    model.invalidate_packed_weights()
    with torch.no_grad():
        changed = model(x, ea, ei, node_types=types)[0][-1]
    assert not torch.equal(changed, good)


def test_umma_selftest_gemm():
    """The tcgen05 building blocks in isolation: D = A . W^T for one 128x64x64 tile."""
    import pgmp_b200._native as nv
    g = torch.Generator().manual_seed(0)
    A = torch.randn(128, 64, generator=g)
    W = torch.randn(64, 64, generator=g)
    D = torch.full((128, 64), float("nan"), device=DEV)
    A_d, W_d = A.to(DEV), W.to(DEV)          # keep the device copies alive across the launch
    nv.check(nv.lib().pgmp_selftest_umma(A_d.data_ptr(), W_d.data_ptr(), D.data_ptr(), nv.current_stream()))
    torch.cuda.synchronize()
    want = (A.double() @ W.double().t()).numpy()
    assert_close(D.cpu().numpy(), want, 1e-4, "bf16x3 GEMM")     # ~2^-16 per product term


def test_umma_selftest_gemm_a_operand_in_tensor_memory():
    """The TS form of tcgen05.mma: A written to tensor memory with tcgen05.st (row = lane, element k in half k & 1 of
    column k / 2), B in shared memory."""
    import pgmp_b200._native as nv
    g = torch.Generator().manual_seed(4)
    A = torch.randn(128, 64, generator=g)
    W = torch.randn(64, 64, generator=g)
    D = torch.full((128, 64), float("nan"), device=DEV)
    A_d, W_d = A.to(DEV), W.to(DEV)
    nv.check(nv.lib().pgmp_selftest_umma_ts(A_d.data_ptr(), W_d.data_ptr(), D.data_ptr(), nv.current_stream()))
    torch.cuda.synchronize()
    assert_close(D.cpu().numpy(), (A.double() @ W.double().t()).numpy(), 1e-4, "bf16x3 GEMM, TS form")


@pytest.mark.parametrize("name", list(MPN_CASES))
def test_mpn_tensor_core_within_logit_tolerance(name):
    """north_star: logits within 1e-3 relative error with fp32-accumulated bf16 tensor-core math."""
    cfg, g, model = mpn_case(name, precision="tc")
    pe, pn, pc, _ = run_mpn(model, g)
    sd = {k: v.cpu().numpy() for k, v in model.state_dict().items()}
    ope, opn, opc = oracle.mpn.node_classification_mpn_forward(sd, cfg, g["x"], g["edge_attr"], g["edge_index"],
                                                               g["joint_det"][:, 2])
    gold = golden("mpn_" + name)
    for kind, got, want in (("edge", pe, ope), ("node", pn, opn), ("class", pc, opc)):
        assert len(got) == len(want)
        for i, (a, b) in enumerate(zip(got, want)):
            a = a.cpu().numpy()
            assert_close(a, b, LOGIT_TOL, f"{name}:{kind}_{i} vs oracle")
            assert_close(a, gold[f"{kind}_{i}"], LOGIT_TOL, f"{name}:{kind}_{i} vs reference")


@pytest.mark.parametrize("over", [
    dict(EDGE_EMB=dict(OUTPUT_SIZES=[32, 96, 64])),                      # a layer wider than 64: SIMT edge embedding + image conversion
    dict(NODE_EMB=dict(OUTPUT_SIZES=[96, 64])),                          # not the 128-128-64-64 chain: SIMT node embedding
    dict(EDGE_CLASS=dict(OUTPUT_SIZES=[48, 1])),                         # not the 64-64-32-1 head: SIMT edge head on the images
    dict(SKIP=False, EDGE_EMB=dict(OUTPUT_SIZES=[80, 64])),              # no C rows at all
    dict(NODE_CLASS=dict(OUTPUT_SIZES=[48, 1])),                         # not the 64-64-32-1 head: SIMT node / class heads
], ids=["wide_edge_emb", "short_node_emb", "short_edge_head", "noskip_wide_edge_emb", "short_node_head"])
def test_mpn_tensor_core_fallback_shapes(over):
    """Layer shapes the tensor-core kernels do not cover run the SIMT stage inside the tensor-core forward; results
    stay within the logit tolerance of the oracle."""
    cfg = pgmp_b200.config.flagship_mpn_config(17, STEPS=3, B200_PRECISION="tc")
    for k, v in over.items():
        if isinstance(v, dict):
            for kk, vv in v.items():
                setattr(getattr(cfg, k), kk, vv)
        else:
            setattr(cfg, k, v)
    data, gcfg, nj = gc_inputs("knn_small")
    g = oracle.gc.construct_graph(data["scoremaps"], data["tagmaps"], data["features"], gcfg, nj, masks=data["masks"])
    model = synthetic.synth_mpn_state_dict(get_mpn_model(cfg), 21).eval().to(DEV)
    pe, pn, pc, _ = run_mpn(model, g)
    sd = {k: v.cpu().numpy() for k, v in model.state_dict().items()}
    ope, opn, opc = oracle.mpn.node_classification_mpn_forward(sd, cfg, g["x"], g["edge_attr"], g["edge_index"],
                                                               g["joint_det"][:, 2])
    for kind, got, want in (("edge", pe, ope), ("node", pn, opn), ("class", pc, opc)):
        assert len(got) == len(want)
        for i, (a, b) in enumerate(zip(got, want)):
            assert_close(a.cpu().numpy(), b, LOGIT_TOL, f"{kind}_{i} vs oracle")


# ------------------------------------------------------------------------------------------------
# grouping tail: threshold -> GAEC multicut -> persons (bit-exact on identical logits)
# ------------------------------------------------------------------------------------------------
def _oracle_graph_for(gc_name):
    data, gcfg, nj = gc_inputs(gc_name)
    return oracle.gc.construct_graph(data["scoremaps"], data["tagmaps"], data["features"], gcfg, nj, masks=data["masks"]), nj


def _check_grouping(g, nj, logits, th, gold=None, cc_method="GAEC"):
    from pgmp_b200.Utils import group_persons
    res = group_persons(torch.from_numpy(g["joint_det"]).to(DEV), torch.from_numpy(logits["node_logits"]).to(DEV),
                        torch.from_numpy(g["edge_index"]).to(DEV), torch.from_numpy(logits["edge_logits"]).to(DEV),
                        torch.from_numpy(logits["class_logits"]).to(DEV), torch.from_numpy(g["batch_index"]).to(DEV),
                        nj, node_threshold=th, cc_method=cc_method)
    assert len(res) == len(np.unique(g["batch_index"]))
    for b, got in enumerate(res):
        sub = synthetic.image_subgraph(g, logits, b)
        want = oracle.grouping.pred_to_person(sub["joint_det"], sub["node_logits"], sub["edge_index"],
                                              sub["edge_logits"], sub["class_logits"], th, nj, cc_method=cc_method)
        if want is None:
            assert got is None
            continue
        persons, mutant, labels = got
        wp, wm, wl = want
        assert np.array_equal(labels.cpu().numpy(), wl), f"image {b}: person labels differ"
        assert persons.shape == wp.shape and bool(mutant) == bool(wm)
        if wp.size:
            assert np.array_equal(persons[:, :, :2], wp[:, :, :2])
            np.testing.assert_allclose(persons[:, :, 2], wp[:, :, 2], rtol=1e-6)
        if gold is not None:
            assert np.array_equal(labels.cpu().numpy(), gold[f"labels_{b}"])
            assert np.array_equal(persons[:, :, :2], gold[f"persons_{b}"][:, :, :2])


@pytest.mark.parametrize("name,gc_name,seed", [("group_threshold_knn_small", "knn_small", 0),
                                               ("group_threshold_crowdpose", "crowdpose", 2)])
def test_grouping_threshold_method_matches_reference(name, gc_name, seed):
    """CC_METHOD = "threshold" (Utils.py:508-509): fixtures come from the reference's own code (no stand-in)."""
    g, nj = _oracle_graph_for(gc_name)
    logits = synthetic.synth_group_logits(g["joint_det"], g["batch_index"], g["edge_index"], num_joints=nj, seed=seed)
    _check_grouping(g, nj, logits, 0.1, gold=golden(name), cc_method="threshold")
    _check_grouping(g, nj, dict(logits, edge_logits=logits["edge_logits"] * 0.3), 0.1, cc_method="threshold")   # few joins


@pytest.mark.parametrize("name,gc_name,seed", [("group_greedy_knn_small", "knn_small", 0), ("group_greedy_fully_small", "fully_small", 1),
                                               ("group_greedy_crowdpose", "crowdpose", 2)])
def test_grouping_greedy_method_matches_reference(name, gc_name, seed):
    """CC_METHOD = "greedy" (greedy_person_construction, Utils.py:517-626): pure numpy in the reference, so the whole
    grouping is pinned end to end -- fixtures from the reference's own code; labels = the core node that claimed each node."""
    g, nj = _oracle_graph_for(gc_name)
    logits = synthetic.synth_group_logits(g["joint_det"], g["batch_index"], g["edge_index"], num_joints=nj, seed=seed)
    _check_grouping(g, nj, logits, 0.1, gold=golden(name), cc_method="greedy")
    _check_grouping(g, nj, dict(logits, node_logits=logits["node_logits"] * 0.2), 0.3, cc_method="greedy")   # scores around 0.5


@pytest.mark.parametrize("name,gc_name,seed", [("group_knn_small", "knn_small", 0), ("group_fully_small", "fully_small", 1),
                                               ("group_crowdpose", "crowdpose", 2)])
def test_grouping_bit_exact_on_identical_logits(name, gc_name, seed):
    g, nj = _oracle_graph_for(gc_name)
    logits = synthetic.synth_group_logits(g["joint_det"], g["batch_index"], g["edge_index"], num_joints=nj, seed=seed)
    _check_grouping(g, nj, logits, 0.1, golden(name))


def test_grouping_edge_cases():
    g, nj = _oracle_graph_for("knn_small")
    logits = synthetic.synth_group_logits(g["joint_det"], g["batch_index"], g["edge_index"], num_joints=nj, seed=5)
    # every node below the threshold in image 1 -> no kept edge -> None (Utils.py:1452,1457)
    lo = dict(logits, node_logits=logits["node_logits"].copy())
    lo["node_logits"][g["batch_index"] == 1] = -20.0
    _check_grouping(g, nj, lo, 0.1)
    # all edges repulsive -> singletons only -> no persons; all attractive -> one person per image
    _check_grouping(g, nj, dict(logits, edge_logits=-np.abs(logits["edge_logits"]) - 0.1), 0.1)
    _check_grouping(g, nj, dict(logits, edge_logits=np.abs(logits["edge_logits"]) + 0.1), 0.5)


def test_full_pipeline_heatmaps_to_persons():
    """construct_graph -> mpn.forward -> group_persons entirely on the device; the grouping must equal the
    oracle's grouping of the same logits."""
    from pgmp_b200.Utils import group_persons
    ret, _ = run_gc("knn_small")
    cfg, g, model = mpn_case("flagship")
    with torch.no_grad():
        pe, pn, pc, _ = model(ret[0], ret[1], ret[2], node_types=ret[7][:, 2])
    res = group_persons(ret[7], pn[-1], ret[2], pe[-1], pc[-1], ret[12], 17, node_threshold=0.5, detector_scores=ret[11])
    logits = dict(node_logits=pn[-1].cpu().numpy(), edge_logits=pe[-1].cpu().numpy(), class_logits=pc[-1].cpu().numpy(),
                  num_joints=17)
    for b, got in enumerate(res):
        sub = synthetic.image_subgraph(g, logits, b)
        want = oracle.grouping.pred_to_person(sub["joint_det"], sub["node_logits"], sub["edge_index"],
                                              sub["edge_logits"], sub["class_logits"], 0.5, 17)
        assert (got is None) == (want is None)
        if want is not None:
            assert np.array_equal(got[2].cpu().numpy(), want[2])
            assert got[0].shape == want[0].shape


# ------------------------------------------------------------------------------------------------
# size-independent properties at the BASELINE workload size (32 images of 17 x 512 x 512, 30 per joint)
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def full_size():
    B, J, S, K = 32, 17, 512, 30
    sm = torch.from_numpy(np.stack([synthetic.synth_scoremap(b, J, S, K) for b in range(B)])).to(DEV)
    gen = torch.Generator(device=DEV).manual_seed(7)
    feat = torch.randn(B, 128, S, S, device=DEV, generator=gen)
    tags = torch.randn(B, J, S, S, device=DEV, generator=gen)
    cfg = pgmp_b200.config.bench_gc_config(k=K, graph_type="knn")
    ret = get_graph_constructor(cfg, scoremaps=sm, tagmaps=tags, features=feat, joints_gt=None, factor_list=None,
                                masks=None, device=DEV, testing=True, heatmaps=None, num_joints=J).construct_graph()
    return dict(B=B, J=J, S=S, K=K, sm=sm, feat=feat, tags=tags, cfg=cfg, ret=ret)


def test_full_size_graph_properties(full_size):
    f = full_size
    x, ea, ei, jd, js, bi, jt = (f["ret"][i] for i in (0, 1, 2, 7, 11, 12, 14))
    B, J, K, S = f["B"], f["J"], f["K"], f["S"]
    N = B * J * K
    assert jd.shape == (N, 3) and x.shape == (N, 128) and ei.shape[0] == 2
    # candidates: exactly K per (image, type), (type, y, x)-sorted inside every image, values are gathers of the inputs
    key = ((bi * J + jd[:, 2]) * S + jd[:, 1]) * S + jd[:, 0]
    assert bool((key[1:] > key[:-1]).all())
    assert bool((torch.bincount(bi * J + jd[:, 2], minlength=B * J) == K).all())
    assert torch.equal(js, f["sm"][bi, jd[:, 2], jd[:, 1], jd[:, 0]])
    assert torch.equal(x, f["feat"][bi, :, jd[:, 1], jd[:, 0]])
    assert torch.equal(jt, f["tags"][bi, jd[:, 2], jd[:, 1], jd[:, 0]])
    # every candidate is a positive 5x5 maximum of its map
    pooled = torch.nn.functional.max_pool2d(f["sm"], 5, 1, 2)
    assert bool((pooled[bi, jd[:, 2], jd[:, 1], jd[:, 0]] == js).all()) and bool((js > 0).all())
    # edge index: (src, dst)-sorted, no self loops, symmetric, inside one image, degree >= 50
    src, dst = ei
    ekey = src * N + dst
    assert bool((ekey[1:] > ekey[:-1]).all()) and bool((src != dst).all())
    assert bool((bi[src] == bi[dst]).all())
    rkey, _ = torch.sort(dst * N + src)
    assert torch.equal(rkey, ekey)
    assert int(torch.bincount(src, minlength=N).min()) >= 50
    # edge attributes are the closed form of CG.py:305-325
    want = torch.zeros_like(ea)
    want[:, 0] = (jd[dst, 0] - jd[src, 0]).float() / S
    want[:, 1] = (jd[dst, 1] - jd[src, 1]).float() / S
    ar = torch.arange(ei.shape[1], device=DEV)
    want[ar, 2 + jd[src, 2]] = 1
    want[ar, 2 + jd[dst, 2]] = 1
    assert torch.equal(ea, want)


def test_full_size_batch_independence_and_idempotence(full_size):
    """Images are independent units: image b's block of the batched graph equals the graph of image b alone
    (with node ids offset); a second run is bit-identical."""
    f = full_size
    J = f["J"]
    again = get_graph_constructor(f["cfg"], scoremaps=f["sm"], tagmaps=f["tags"], features=f["feat"], joints_gt=None,
                                  factor_list=None, masks=None, device=DEV, testing=True, heatmaps=None,
                                  num_joints=J).construct_graph()
    for i in (0, 1, 2, 7, 11, 12, 14):
        assert torch.equal(again[i], f["ret"][i])
    b = 17
    one = get_graph_constructor(f["cfg"], scoremaps=f["sm"][b:b + 1], tagmaps=f["tags"][b:b + 1], features=f["feat"][b:b + 1],
                                joints_gt=None, factor_list=None, masks=None, device=DEV, testing=True, heatmaps=None,
                                num_joints=J).construct_graph()
    bi = f["ret"][12]
    nsel = bi == b
    off = int(torch.nonzero(nsel)[0])
    esel = nsel[f["ret"][2][0]]
    assert torch.equal(one[7], f["ret"][7][nsel]) and torch.equal(one[0], f["ret"][0][nsel])
    assert torch.equal(one[2] + off, f["ret"][2][:, esel]) and torch.equal(one[1], f["ret"][1][esel])


@pytest.mark.parametrize("precision", ["tc", "fp32"])
def test_full_size_mpn_batch_independence(full_size, precision):
    """The MPN has no cross-image coupling in eval mode: logits of image b inside the 32-image batch equal the
    logits of image b run alone (same kernels, different tiling) within tolerance; runs are deterministic."""
    f = full_size
    J = f["J"]
    cfg = pgmp_b200.config.flagship_mpn_config(J, B200_PRECISION=precision)
    model = synthetic.synth_mpn_state_dict(get_mpn_model(cfg), 11).eval().to(DEV)
    ret = f["ret"]
    with torch.no_grad():
        pe, pn, pc, _ = model(ret[0], ret[1], ret[2], node_types=ret[7][:, 2])
        pe2, pn2, pc2, _ = model(ret[0], ret[1], ret[2], node_types=ret[7][:, 2])
    assert torch.equal(pe[-1], pe2[-1]) and torch.equal(pn[-1], pn2[-1]) and torch.equal(pc[-1], pc2[-1])
    assert bool(torch.isfinite(pe[-1]).all()) and bool(torch.isfinite(pc[-1]).all())
    b = 5
    nsel = ret[12] == b
    off = int(torch.nonzero(nsel)[0])
    esel = nsel[ret[2][0]]
    with torch.no_grad():
        qe, qn, qc, _ = model(ret[0][nsel], ret[1][esel], ret[2][:, esel] - off, node_types=ret[7][nsel, 2])
    tol = LOGIT_TOL if precision == "tc" else FP32_TOL
    assert_close(qe[-1].cpu().numpy(), pe[-1][esel].cpu().numpy(), tol, "edge logits, image alone vs in batch")
    assert_close(qn[-1].cpu().numpy(), pn[-1][nsel].cpu().numpy(), tol, "node logits")
    assert_close(qc[-1].cpu().numpy(), pc[-1][nsel].cpu().numpy(), tol, "class logits")


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs[2..3] at full per-image size: w48 / 640 px fully connected, CrowdPose-shaped (14 joints,
# 60 candidates per joint, kNN and fully connected: 704 760 edges per image)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["w48_640_fully", "crowdpose_knn", "crowdpose_fully"])
def test_other_configs_full_image_size(name):
    """BASELINE configs[2..3] at full per-image size and the full 10 steps against the UNMODIFIED reference's outputs
    (``tests/golden/full_*.npz``, written by ``make_golden_fullsize.py``): graph-constructor outputs bit-exact (digests),
    fp32 and tensor-core logits of image 0 within tolerance -- every node / class logit, the edge logits at 8192 fixed
    positions and the l2 norm over all edges.  Image 0 sits in a batch of two: graph invariants over the batch, and
    image 1 alone equals its block of the batch."""
    inp_kw, cfg_over, mpn_over, seed = FULL_CASES[name]
    gold = golden("full_" + name)
    J, S = inp_kw["num_joints"], inp_kw["size"]
    K, graph = cfg_over["k"], cfg_over["graph_type"]
    B = 2
    data = synthetic.synth_batch(**dict(inp_kw, batch=B))
    sm, feat, tags = (torch.from_numpy(data[k]).to(DEV) for k in ("scoremaps", "features", "tagmaps"))
    gcfg = gc_config_for(pgmp_b200.config, cfg_over)

    def gc(sl):
        return get_graph_constructor(gcfg, scoremaps=sm[sl], tagmaps=tags[sl], features=feat[sl], joints_gt=None,
                                     factor_list=None, masks=None, device=DEV, testing=True, heatmaps=None,
                                     num_joints=J).construct_graph()
    ret = gc(slice(0, B))
    x, ea, ei, jd, bi = ret[0], ret[1], ret[2], ret[7], ret[12]
    n = J * K
    N = B * n
    assert jd.shape == (N, 3) and bool((torch.bincount(bi * J + jd[:, 2], minlength=B * J) == K).all())
    src, dst = ei
    ekey = src * N + dst
    assert bool((ekey[1:] > ekey[:-1]).all()) and bool((src != dst).all()) and bool((bi[src] == bi[dst]).all())
    if graph == "fully":
        assert ei.shape[1] == B * n * (n - 1)
    else:
        rkey, _ = torch.sort(dst * N + src)
        assert torch.equal(rkey, ekey) and int(torch.bincount(src, minlength=N).min()) >= 50
    assert ea.shape == (ei.shape[1], J + 2) and bool((ea[:, 2:].sum(1) >= 1).all())
    # image 0's block == the reference's graph of that image (node ids of image 0 carry no offset)
    n0 = bi == 0
    e0 = n0[src]
    block = dict(x=x[n0], edge_attr=ea[e0], edge_index=ei[:, e0], joint_det=jd[n0], joint_scores=ret[11][n0],
                 batch_index=bi[n0], joint_tags=ret[14][n0])
    for k, v in block.items():
        assert_matches_golden(gold, k, v.contiguous().cpu().numpy())
    idx = torch.from_numpy(full_edge_sample(int(e0.sum()), name)).to(DEV)
    for prec in ("fp32", "tc"):
        mcfg = mpn_config_for(pgmp_b200.config, "flagship_mpn_config", dict(mpn_over, B200_PRECISION=prec))
        assert mcfg.STEPS == 10
        model = synthetic.synth_mpn_state_dict(get_mpn_model(mcfg), seed).eval().to(DEV)
        with torch.no_grad():
            pe, pn, pc, _ = model(x, ea, ei, node_types=jd[:, 2])
        assert len(pe) == int(gold["n_edge"]) and len(pn) == int(gold["n_node"])
        tol = LOGIT_TOL if prec == "tc" else FP32_TOL_FULL
        edge0 = pe[-1][e0]
        assert_close(edge0[idx].cpu().numpy(), gold["edge_sample"], tol, f"{name}/{prec}: sampled edge logits vs reference")
        l2 = float(torch.linalg.vector_norm(edge0.double()))
        assert abs(l2 - float(gold["edge_l2"])) <= tol * float(gold["edge_l2"]), f"{name}/{prec}: l2 norm of all edge logits"
        assert_close(pn[-1][n0].cpu().numpy(), gold["node"], tol, f"{name}/{prec}: node logits vs reference")
        assert_close(pc[-1][n0].cpu().numpy(), gold["cls"], tol, f"{name}/{prec}: class logits vs reference")
        if prec == "tc":          # image 1 alone == its block of the batch
            one = gc(slice(1, 2))
            with torch.no_grad():
                qe, qn, qc, _ = model(one[0], one[1], one[2], node_types=one[7][:, 2])
            nsel = bi == 1
            esel = nsel[src]
            assert_close(qe[-1].cpu().numpy(), pe[-1][esel].cpu().numpy(), LOGIT_TOL, f"{name}: edge logits alone vs in batch")
            assert_close(qn[-1].cpu().numpy(), pn[-1][nsel].cpu().numpy(), LOGIT_TOL, f"{name}: node logits alone vs in batch")


@pytest.mark.parametrize("precision,aggr_sub", [("tc", "node_edge_attn"), ("tc", "None"), ("fp32", "node_edge_attn")])
def test_mpn_bins_larger_than_a_tile(precision, aggr_sub):
    """150 nodes of one type in a complete graph: every (target, type 0) bin holds ~150 edges, i.e. runs that cover
    whole warps, straddle 128-slot tiles and have three parts -- the slow paths of the run reduction and of the node
    update -- against the oracle."""
    rng = np.random.default_rng(5)
    n = 200
    types = np.concatenate([np.zeros(150, np.int64), rng.integers(1, 17, size=n - 150)])
    rng.shuffle(types)
    xy = rng.integers(0, 512, size=(n, 2))
    src, dst = np.nonzero(~np.eye(n, dtype=bool))                       # all ordered pairs, (src, dst)-sorted
    ea = np.zeros((src.size, 19), np.float32)
    ea[:, 0] = (xy[dst, 0] - xy[src, 0]) / 512.0
    ea[:, 1] = (xy[dst, 1] - xy[src, 1]) / 512.0
    ea[np.arange(src.size), 2 + types[src]] = 1
    ea[np.arange(src.size), 2 + types[dst]] = 1
    x = rng.standard_normal((n, 128)).astype(np.float32)
    ei = np.stack([src, dst]).astype(np.int64)
    cfg = pgmp_b200.config.flagship_mpn_config(17, STEPS=3, AGGR_SUB=aggr_sub, B200_PRECISION=precision)
    model = synthetic.synth_mpn_state_dict(get_mpn_model(cfg), 31).eval().to(DEV)
    with torch.no_grad():
        pe, pn, pc, _ = model(torch.from_numpy(x).to(DEV), torch.from_numpy(ea).to(DEV), torch.from_numpy(ei).to(DEV),
                              node_types=torch.from_numpy(types).to(DEV))
    sd = {k: v.cpu().numpy() for k, v in model.state_dict().items()}
    ope, opn, opc = oracle.mpn.node_classification_mpn_forward(sd, cfg, x, ea, ei, types)
    tol = LOGIT_TOL if precision == "tc" else FP32_TOL
    for kind, got, want in (("edge", pe, ope), ("node", pn, opn), ("class", pc, opc)):
        for i, (a, b) in enumerate(zip(got, want)):
            assert_close(a.cpu().numpy(), b, tol, f"{kind}_{i} vs oracle")
