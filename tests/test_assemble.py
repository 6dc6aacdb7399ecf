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
