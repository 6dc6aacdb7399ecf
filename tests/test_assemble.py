"""Scoremap assembly in front of the NMS (hrnet.py:587-611): the numpy oracle against torch's own
``interpolate`` on the CPU (the reference's operation), and the CUDA kernel against the oracle (bit-exact: the same
separately rounded products and sums)."""
import numpy as np
import pytest
import torch

import oracle.assemble as A

SHAPES = [(2, 5, 16, 24, 32, 48), (1, 3, 7, 9, 14, 18), (1, 2, 8, 8, 20, 13), (2, 17, 64, 64, 128, 128)]


def make(B, J, h, w, H, W, seed=0):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((B, 2 * J, h, w)).astype(np.float32), rng.standard_normal((B, J, H, W)).astype(np.float32)


@pytest.mark.parametrize("shape", SHAPES)
def test_oracle_matches_torch_interpolate(shape):
    B, J, h, w, H, W = shape
    s1, s2 = make(*shape)
    up = torch.nn.functional.interpolate(torch.from_numpy(s1), size=(H, W), mode="bilinear", align_corners=False)
    want_score = ((torch.from_numpy(s2) + up[:, :J]) / 2).numpy()
    score, tags = A.hr_process_output(s1, s2, J, "avg")
    assert np.abs(score - want_score).max() <= 1e-6 * max(1.0, np.abs(want_score).max())
    assert np.abs(tags - up[:, J:].numpy()).max() <= 1e-6 * max(1.0, np.abs(s1).max())
    small, _ = A.hr_process_output(s1, s2, J, "small")
    assert np.abs(small - up[:, :J].numpy()).max() <= 1e-6 * max(1.0, np.abs(s1).max())


@pytest.mark.gpu
@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("mode", ["avg", "small", "large"])
def test_cuda_assembly_bit_exact_with_oracle(shape, mode):
    from pgmp_b200.graph_constructor import hr_process_output
    B, J, h, w, H, W = shape
    s1, s2 = make(*shape, seed=1)
    feats = object()
    score, f, tags = hr_process_output(((torch.from_numpy(s1).cuda(), torch.from_numpy(s2).cuda()), feats), mode, J)
    want_score, want_tags = A.hr_process_output(s1, s2, J, mode)
    assert f is feats
    assert np.array_equal(score.cpu().numpy(), want_score)
    assert np.array_equal(tags.cpu().numpy(), want_tags)


@pytest.mark.gpu
def test_cuda_assembly_full_size_properties():
    """32 x 17 x 512 x 512: the average of a stage with its own down-sampled copy keeps constants, and the kernel's
    output feeds the graph constructor."""
    from pgmp_b200.graph_constructor import hr_process_output
    B, J, H = 4, 17, 512
    s2 = torch.rand(B, J, H, H, device="cuda")
    s1 = torch.full((B, 2 * J, H // 2, H // 2), 0.25, device="cuda")
    score, _, tags = hr_process_output(((s1, s2), None), "avg", J)
    assert torch.equal(score, (s2 + 0.25) * 0.5) and bool((tags == 0.25).all())


@pytest.mark.gpu
def test_assembled_scoremaps_feed_the_graph_constructor():
    """Both stages of the head -> assembly kernel -> graph constructor, against the oracle chain on the same inputs:
    candidates, scores and edge index bit-exact (the assembled maps are bit-identical, so is everything downstream)."""
    import oracle
    import pgmp_b200
    import pgmp_b200.synthetic as synthetic
    from pgmp_b200.graph_constructor import get_graph_constructor, hr_process_output
    J, K, S = 17, 10, 128
    data = synthetic.synth_batch(2, J, S, K, persons=3)
    rng = np.random.default_rng(3)
    s2 = data["scoremaps"]
    # a half-resolution stage: 2x2 block means of the full-resolution maps + noise, and random tag maps
    s1_heat = s2.reshape(2, J, S // 2, 2, S // 2, 2).mean((3, 5)).astype(np.float32) + rng.uniform(0, 0.01, (2, J, S // 2, S // 2)).astype(np.float32)
    s1 = np.concatenate([s1_heat, rng.standard_normal((2, J, S // 2, S // 2)).astype(np.float32)], 1)
    cfg = pgmp_b200.config.bench_gc_config(k=K, graph_type="knn")
    score, feats, tags = hr_process_output(((torch.from_numpy(s1).cuda(), torch.from_numpy(s2).cuda()), torch.from_numpy(data["features"]).cuda()), "avg", J)
    ret = get_graph_constructor(cfg, scoremaps=score, tagmaps=tags, features=feats, joints_gt=None, factor_list=None, masks=None,
                                device="cuda:0", testing=True, heatmaps=None, num_joints=J).construct_graph()
    o_score, o_tags = oracle.assemble.hr_process_output(s1, s2, J, "avg")
    want = oracle.gc.construct_graph(o_score, o_tags, data["features"], cfg, J)
    for key, slot in (("x", 0), ("edge_attr", 1), ("edge_index", 2), ("joint_det", 7), ("joint_scores", 11), ("joint_tags", 14)):
        assert np.array_equal(ret[slot].cpu().numpy(), want[key]), key


COCO_FLIP = [0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15]      # FLIP_CONFIG['COCO'] (hr_utils)


def _stages(B, J, S, K, seed, persons=3):
    """Full-resolution stage = synthetic heatmaps, half-resolution stage = 2 x 2 block means + noise, random tag maps."""
    import pgmp_b200.synthetic as synthetic
    data = synthetic.synth_batch(B, J, S, K, persons=persons)
    rng = np.random.default_rng(seed)
    s2 = data["scoremaps"]
    heat = s2.reshape(B, J, S // 2, 2, S // 2, 2).mean((3, 5)).astype(np.float32) + rng.uniform(0, 0.01, (B, J, S // 2, S // 2)).astype(np.float32)
    s1 = np.concatenate([heat, rng.standard_normal((B, J, S // 2, S // 2)).astype(np.float32)], 1)
    return data, s1, s2


def test_oracle_flip_average_matches_the_reference_operations():
    """oracle.assemble.flip_average against the reference's own torch operations (PoseEstimation.py:377-402,
    multi_scales_testing.py:162) on random maps."""
    rng = np.random.default_rng(5)
    a, af = rng.standard_normal((2, 17, 12, 20)).astype(np.float32), rng.standard_normal((2, 17, 12, 20)).astype(np.float32)
    heat_flip = torch.flip(torch.from_numpy(af), [3])[:, COCO_FLIP, :, :]
    want = ((torch.from_numpy(a) + heat_flip) / 2.0).numpy()
    assert np.array_equal(A.flip_average(a, af, COCO_FLIP), want)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["avg", "small", "avg_keep", "avg_flip", "avg_flip_keep", "avg_mask", "no_threshold", "pool3_rect",
                                  "generic_scale", "generic_scale_flip"])
def test_fused_assembly_in_the_nms_loader_is_bit_exact(case):
    """HeadStages as the ``scoremaps`` argument: the NMS loader warps evaluate the assembled map row by row into the
    shared-memory ring (pgmp_gc_detect_fused); every output of construct_graph() must equal the oracle chain
    (assembly -> graph constructor) bit for bit, with and without the materialised map."""
    import oracle
    import pgmp_b200
    from pgmp_b200.graph_constructor import HeadStages, get_graph_constructor
    J, K = 17, 10
    B, S = (2, 128)
    data, s1, s2 = _stages(B, J, S, K, seed=3)
    mode = "small" if case == "small" else "avg"
    if case == "pool3_rect":          # non-square maps, another pool kernel
        s1, s2 = np.ascontiguousarray(s1[:, :, :40, :]), np.ascontiguousarray(s2[:, :, :80, :])
        data = {k: (np.ascontiguousarray(v[:, :, :80, :]) if v.ndim == 4 and v.shape[2] == S else v) for k, v in data.items()}
    if case.startswith("generic_scale"):   # not an exact doubling: the per-column index / weight tables
        rng = np.random.default_rng(8)
        s1 = np.concatenate([rng.uniform(0, 0.02, (B, J, 48, 44)), rng.standard_normal((B, J, 48, 44))], 1).astype(np.float32)
    kw = {}
    if case == "no_threshold":
        kw["DETECT_THRESHOLD"] = 2.0
    if case == "pool3_rect":
        kw["POOL_KERNEL_SIZE"] = 3
    masks = None
    if case == "avg_mask":
        kw["MASK_CROWDS"] = True
        masks = (np.random.default_rng(1).uniform(0, 1, (B, S, S)) > 0.2).astype(np.float32)
    cfg = pgmp_b200.config.bench_gc_config(k=K, graph_type="knn", **kw)
    t = lambda a: torch.from_numpy(a).cuda()
    o_score, o_tags = oracle.assemble.hr_process_output(s1, s2, J, mode)
    flipped = flip_index = None
    if "flip" in case:
        _, s1f, s2f = _stages(B, J, S, K, seed=11)
        if case.startswith("generic_scale"):
            s1f = np.random.default_rng(9).uniform(0, 0.02, s1.shape).astype(np.float32)
        s1f, s2f = np.ascontiguousarray(s1f[..., ::-1]), np.ascontiguousarray(s2f[..., ::-1])
        flipped, flip_index = (t(s1f), t(s2f)), COCO_FLIP
        o_flip, _ = oracle.assemble.hr_process_output(s1f, s2f, J, mode)
        o_score = oracle.assemble.flip_average(o_score, o_flip, COCO_FLIP)
    stages = HeadStages((t(s1), t(s2)), J, mode=mode, flipped=flipped, flip_index=flip_index, keep_scoremaps="keep" in case)
    # odd cases: the detections' tags straight from the half-resolution stage, even ones: from the up-sampled maps
    lazy_tags = case in ("avg", "avg_flip", "pool3_rect", "generic_scale")
    ret = get_graph_constructor(cfg, scoremaps=stages, tagmaps=stages if lazy_tags else t(o_tags), features=t(data["features"]), joints_gt=None,
                                factor_list=None, masks=t(masks) if masks is not None else None, device="cuda:0", testing=True,
                                heatmaps=None, num_joints=J).construct_graph()
    want = oracle.gc.construct_graph(o_score, o_tags, data["features"], cfg, J, masks=masks)
    assert want["joint_det"].shape[0] > 50
    for key, slot in (("x", 0), ("edge_attr", 1), ("edge_index", 2), ("joint_det", 7), ("joint_scores", 11), ("joint_tags", 14)):
        assert np.array_equal(ret[slot].cpu().numpy(), want[key]), key
    if "keep" in case or case == "no_threshold":
        assert np.array_equal(stages.scoremaps.cpu().numpy(), o_score)         # the materialised map, bit for bit
    else:
        assert stages.scoremaps is None


@pytest.mark.gpu
def test_fused_assembly_rejects_what_it_does_not_cover():
    import pgmp_b200
    from pgmp_b200.graph_constructor import HeadStages, get_graph_constructor
    J = 4
    s1, s2 = torch.rand(1, 2 * J, 16, 15, device="cuda"), torch.rand(1, J, 32, 30, device="cuda")       # width % 4 != 0
    cfg = pgmp_b200.config.bench_gc_config(k=5, graph_type="knn")
    with pytest.raises(RuntimeError, match="width"):
        get_graph_constructor(cfg, scoremaps=HeadStages((s1, s2), J), tagmaps=None, features=None, joints_gt=None, factor_list=None,
                              masks=None, device="cuda:0", testing=True, heatmaps=None, num_joints=J).construct_graph()
    with pytest.raises(NotImplementedError):
        HeadStages((s1, s2), J, mode="large")
    with pytest.raises(ValueError):
        HeadStages((s1, s2), J, flipped=(s1, s2))


@pytest.mark.gpu
def test_fused_assembly_full_size_equals_the_two_kernel_path():
    """BASELINE size (32 x 17 x 512 x 512, flip-test average on): fused detection = assembly kernels + plain detection."""
    import pgmp_b200
    from pgmp_b200.graph_constructor import HeadStages, get_graph_constructor, hr_process_output
    J, K, B, S = 17, 30, 32, 512
    data, s1, s2 = _stages(B, J, S, K, seed=2, persons=6)
    t = lambda a: torch.from_numpy(a).cuda()
    s1d, s2d = t(s1), t(s2)
    s1f, s2f = torch.flip(s1d, [3]).contiguous(), torch.flip(s2d, [3]).contiguous()
    cfg = pgmp_b200.config.bench_gc_config(k=K, graph_type="knn")
    feat = torch.zeros(B, 4, S, S, device="cuda")
    a0, _, tags = hr_process_output(((s1d, s2d), None), "avg", J)
    a1, _, _ = hr_process_output(((s1f, s2f), None), "avg", J)
    score = (a0 + torch.flip(a1, [3])[:, COCO_FLIP]) / 2.0
    run = lambda sm: get_graph_constructor(cfg, scoremaps=sm, tagmaps=tags, features=feat, joints_gt=None, factor_list=None, masks=None,
                                           device="cuda:0", testing=True, heatmaps=None, num_joints=J).construct_graph()
    want = run(score)
    got = run(HeadStages((s1d, s2d), J, flipped=(s1f, s2f), flip_index=COCO_FLIP))
    assert want[7].shape[0] > 10000
    for slot in (0, 1, 2, 7, 11, 12, 14):
        assert torch.equal(got[slot], want[slot]), slot


@pytest.mark.gpu
def test_fused_assembly_through_the_pipelined_api():
    """GroupingPipeline with HeadStages batches (the detection half, assembly included, runs one batch ahead on the side
    stream) = the serial two-kernel calls, bit for bit."""
    import pgmp_b200
    import pgmp_b200.synthetic as synthetic
    from pgmp_b200.graph_constructor import HeadStages, get_graph_constructor, hr_process_output
    from pgmp_b200.Models.MessagePassingNetwork import get_mpn_model
    from pgmp_b200.pipeline import GroupingPipeline
    J, K, B, S = 17, 10, 2, 128
    cfg = pgmp_b200.config.bench_gc_config(k=K, graph_type="knn")
    mcfg = pgmp_b200.config.flagship_mpn_config(J, STEPS=2, B200_PRECISION="tc")
    model = synthetic.synth_mpn_state_dict(get_mpn_model(mcfg), 1).eval().cuda()
    t = lambda a: torch.from_numpy(a).cuda()
    batches, want = [], []
    for seed in (3, 4, 5):
        data, s1, s2 = _stages(B, J, S, K, seed=seed)
        s1d, s2d, feat = t(s1), t(s2), t(data["features"])
        st = HeadStages((s1d, s2d), J)
        batches.append({"scoremaps": st, "tagmaps": st, "features": feat})
        score, _, tags = hr_process_output(((s1d, s2d), None), "avg", J)
        ret = get_graph_constructor(cfg, scoremaps=score, tagmaps=tags, features=feat, joints_gt=None, factor_list=None, masks=None,
                                    device="cuda:0", testing=True, heatmaps=None, num_joints=J).construct_graph()
        with torch.no_grad():
            pe, pn, pc, _ = model(ret[0], ret[1], ret[2], node_types=ret[7][:, 2])
        want.append((ret, pe[-1].clone(), pn[-1].clone(), pc[-1].clone()))
    outs = list(GroupingPipeline(cfg, model, J, "cuda:0").run(batches))
    assert len(outs) == 3
    for (ret, (pe, pn, pc)), (wret, wpe, wpn, wpc) in zip(outs, want):
        for slot in (0, 1, 2, 7, 11, 12, 14):
            assert torch.equal(ret[slot], wret[slot]), slot
        assert torch.equal(pe[-1], wpe) and torch.equal(pn[-1], wpn) and torch.equal(pc[-1], wpc)


def test_head_stages_host_logic_without_a_gpu():
    """No CPU fallback: CPU stages are rejected when the stand-in is built; argument checks do not need a device."""
    from pgmp_b200.graph_constructor import HeadStages
    s1, s2 = torch.zeros(1, 8, 8, 8), torch.zeros(1, 4, 16, 16)
    with pytest.raises(Exception, match="(?i)cuda"):
        HeadStages((s1, s2), 4)
    with pytest.raises(NotImplementedError):
        HeadStages((s1, s2), 4, mode="large")
    with pytest.raises(ValueError):
        HeadStages((s1, s2), 4, flipped=(s1, s2))
