"""node features = interpolate(feature_gather(feat)) evaluated at the candidates only.

CPU: the oracle restatement against the reference's own operations (``nn.Conv2d(Cin, 128, 3, 1, 1)`` +
``interpolate(bilinear, align_corners=False)`` + the gather of ConstructGraph.py:265,269, run with torch on the host).
GPU: the CUDA kernel behind ``ConvUpsampleFeatures`` against the oracle and against the unfused torch pipeline.
"""
import numpy as np
import pytest
import torch

import pgmp_b200
import pgmp_b200.synthetic as synthetic
import oracle.feature_gather as fg
from helpers import FP32_TOL, assert_close

DEV = "cuda:0"


def _case(seed, B, cin, h, w, cout, H, W, n):
    rng = np.random.default_rng(seed)
    feat = rng.standard_normal((B, cin, h, w)).astype(np.float32)
    wt = (rng.standard_normal((cout, cin, 3, 3)) * 0.1).astype(np.float32)
    bias = rng.standard_normal(cout).astype(np.float32)
    jd = np.stack([rng.integers(0, W, n), rng.integers(0, H, n), rng.integers(0, 17, n)], 1).astype(np.int64)
    jd[:4, 0] = [0, W - 1, 0, W - 1]                    # the four corners: clamped taps + zero padding
    jd[:4, 1] = [0, 0, H - 1, H - 1]
    bi = np.sort(rng.integers(0, B, n)).astype(np.int64)
    return feat, wt, bias, jd, bi


def _reference(feat, wt, bias, jd, bi, H, W, device="cpu"):
    """The reference's operations, verbatim (PoseEstimation.py:79, 442-450; ConstructGraph.py:265)."""
    with torch.no_grad():
        f = torch.nn.functional.conv2d(torch.from_numpy(feat).to(device), torch.from_numpy(wt).to(device),
                                       torch.from_numpy(bias).to(device), 1, 1)
        if (H, W) != tuple(f.shape[2:]):
            f = torch.nn.functional.interpolate(f, size=(H, W), mode="bilinear", align_corners=False)
        jd_t, bi_t = torch.from_numpy(jd).to(device), torch.from_numpy(bi).to(device)
        return f[bi_t, :, jd_t[:, 1], jd_t[:, 0]].cpu().numpy()


@pytest.mark.parametrize("shape", [(2, 32, 24, 20, 128, 48, 40), (1, 8, 16, 16, 128, 64, 64), (2, 32, 20, 28, 64, 20, 28),
                                   (1, 5, 7, 9, 16, 21, 18)], ids=["x2", "x4", "same_size", "x3_odd"])
def test_oracle_matches_reference_operations(shape):
    B, cin, h, w, cout, H, W = shape
    feat, wt, bias, jd, bi = _case(1, B, cin, h, w, cout, H, W, 70)
    ref = _reference(feat, wt, bias, jd, bi, H, W)
    assert_close(fg.features_at_candidates_dense(feat, wt, bias, jd, bi, H, W), ref, FP32_TOL, "dense restatement")
    assert_close(fg.features_at_candidates(feat, wt, bias, jd, bi, H, W), ref, FP32_TOL, "candidate-only restatement")


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 32, 24, 20, 128, 48, 40), (1, 8, 16, 16, 128, 64, 64), (2, 32, 20, 28, 64, 20, 28),
                                   (1, 5, 7, 9, 16, 21, 18), (1, 96, 12, 12, 128, 24, 24)],
                         ids=["x2", "x4", "same_size", "x3_odd", "weights_not_resident"])
def test_kernel_matches_oracle(shape):
    import pgmp_b200._native as nv
    B, cin, h, w, cout, H, W = shape
    feat, wt, bias, jd, bi = _case(2, B, cin, h, w, cout, H, W, 93)
    want = fg.features_at_candidates(feat, wt, bias, jd, bi, H, W)
    f = torch.from_numpy(feat).to(DEV)
    wt_t = torch.from_numpy(wt).to(DEV).permute(2, 3, 1, 0).reshape(-1, cout).contiguous()
    b_t, jd_t, bi_t = torch.from_numpy(bias).to(DEV), torch.from_numpy(jd).to(DEV), torch.from_numpy(bi).to(DEV)
    for fm in (f, f.contiguous(memory_format=torch.channels_last), torch.from_numpy(feat).pin_memory()):
        x = torch.full((jd.shape[0], cout), float("nan"), device=DEV)
        p = nv.GatherConvParams(features=fm.data_ptr(), feat_stride_b=fm.stride(0), feat_stride_c=fm.stride(1),
                                feat_stride_y=fm.stride(2), feat_stride_x=fm.stride(3), cin=cin, height=h, width=w, cout=cout,
                                out_height=H, out_width=W, weight_t=wt_t.data_ptr(), bias=b_t.data_ptr(),
                                joint_det=jd_t.data_ptr(), batch_index=bi_t.data_ptr(), num_nodes=jd.shape[0], x=x.data_ptr())
        nv.check(nv.lib().pgmp_gc_gather_conv(p, nv.current_stream()))
        torch.cuda.synchronize()
        assert_close(x.cpu().numpy(), want, FP32_TOL, "kernel vs oracle")
        assert_close(x.cpu().numpy(), _reference(feat, wt, bias, jd, bi, H, W), FP32_TOL, "kernel vs reference operations")


@pytest.mark.gpu
def test_graph_constructor_with_lazy_features_equals_materialised_features():
    """w32 shapes: 32-channel backbone map at half resolution, 128 node channels, 512-pixel heatmaps.  Everything but x
    is bit-identical to the run on the materialised maps; x agrees to fp32 rounding."""
    from pgmp_b200.graph_constructor import ConvUpsampleFeatures, get_graph_constructor
    B, J, S, K = 3, 17, 512, 30
    sm = torch.from_numpy(np.stack([synthetic.synth_scoremap(b, J, S, K) for b in range(B)])).to(DEV)
    gen = torch.Generator(device=DEV).manual_seed(11)
    backbone = torch.randn(B, 32, S // 2, S // 2, device=DEV, generator=gen)
    tags = torch.randn(B, J, S, S, device=DEV, generator=gen)
    conv = torch.nn.Conv2d(32, 128, 3, 1, 1).to(DEV)
    gcfg = pgmp_b200.config.bench_gc_config(k=K, graph_type="knn")

    def run(features):
        return get_graph_constructor(gcfg, scoremaps=sm, tagmaps=tags, features=features, joints_gt=None, factor_list=None,
                                     masks=None, device=DEV, testing=True, heatmaps=None, num_joints=J).construct_graph()
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False            # cuDNN's default TF32 convolution is ~1e-3 off the fp32 result
    try:
        with torch.no_grad():
            full = torch.nn.functional.interpolate(conv(backbone), size=(S, S), mode="bilinear", align_corners=False)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    want = run(full)
    for fm in (backbone, backbone.cpu().pin_memory()):
        got = run(ConvUpsampleFeatures(fm, conv, (S, S)))
        for i in (1, 2, 7, 11, 12, 14):
            assert torch.equal(got[i], want[i]), i
        assert got[0].shape == want[0].shape
        assert got[0].requires_grad                       # the convolution's parameters require gradients: x carries them on
        assert_close(got[0].detach().cpu().numpy(), want[0].cpu().numpy(), FP32_TOL, "x: lazy vs materialised features")


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 8, 16, 16, 16, 32, 32), (2, 32, 20, 28, 64, 20, 28), (1, 5, 7, 9, 16, 21, 18)],
                         ids=["x2", "same_size", "x3_odd"])
def test_lazy_features_reverse_pass_matches_torch_autograd(shape):
    """Training end to end through ConvUpsampleFeatures: gradients of the backbone map, the feature_gather weight and bias
    from the native reverse pass (_GatherConvFeatures) against torch autograd through the reference's own operations
    (conv2d + interpolate + gather, fp32, TF32 off); overlapping neighbourhoods, image corners, duplicate pixels; the
    reverse pass is bit-reproducible (no atomics)."""
    from pgmp_b200.graph_constructor import _GatherConvFeatures
    import pgmp_b200._native as nv
    B, cin, h, w, cout, H, W = shape
    feat, wt, bias, jd, bi = _case(5, B, cin, h, w, cout, H, W, 77)
    jd[5:9] = jd[4]                                           # several candidates at one pixel
    coeff = np.random.default_rng(3).standard_normal((jd.shape[0], cout)).astype(np.float32)
    t = lambda a: torch.from_numpy(a).to(DEV)
    jd_t, bi_t, c_t = t(jd), t(bi), t(coeff)

    tf32c, tf32m = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        f0, w0, b0 = t(feat).requires_grad_(True), t(wt).requires_grad_(True), t(bias).requires_grad_(True)
        full = torch.nn.functional.conv2d(f0, w0, b0, 1, 1)
        if (H, W) != (h, w):
            full = torch.nn.functional.interpolate(full, size=(H, W), mode="bilinear", align_corners=False)
        (full[bi_t, :, jd_t[:, 1], jd_t[:, 0]] * c_t).sum().backward()

        grads = []
        for _ in range(2):
            f1, w1, b1 = t(feat).requires_grad_(True), t(wt).requires_grad_(True), t(bias).requires_grad_(True)
            x = torch.empty((jd.shape[0], cout), device=DEV)
            wt_t = w1.detach().permute(2, 3, 1, 0).reshape(-1, cout).contiguous()
            p = nv.GatherConvParams(features=f1.data_ptr(), feat_stride_b=f1.stride(0), feat_stride_c=f1.stride(1),
                                    feat_stride_y=f1.stride(2), feat_stride_x=f1.stride(3), cin=cin, height=h, width=w, cout=cout,
                                    out_height=H, out_width=W, weight_t=wt_t.data_ptr(), bias=b1.data_ptr(),
                                    joint_det=jd_t.data_ptr(), batch_index=bi_t.data_ptr(), num_nodes=jd.shape[0], x=x.data_ptr())
            nv.check(nv.lib().pgmp_gc_gather_conv(p, nv.current_stream()))
            (_GatherConvFeatures.apply(f1, w1, b1, x, bi_t, jd_t, (H, W)) * c_t).sum().backward()
            grads.append((f1.grad.clone(), w1.grad.clone(), b1.grad.clone()))
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32c, tf32m
    for got, want, name in zip(grads[0], (f0.grad, w0.grad, b0.grad), ("d feat", "d weight", "d bias")):
        assert got.shape == want.shape
        assert_close(got.cpu().numpy(), want.cpu().numpy(), 2e-5, name + ": native reverse pass vs torch autograd")
    for a, b in zip(grads[0], grads[1]):
        assert torch.equal(a, b)                             # fixed summation order


@pytest.mark.gpu
def test_graph_constructor_trains_through_lazy_features():
    """construct_graph() with ConvUpsampleFeatures whose parameters require gradients returns an x that back-propagates
    into the backbone map and the convolution (PoseEstimation.py:64-66 under train.py:232)."""
    from pgmp_b200.graph_constructor import ConvUpsampleFeatures, get_graph_constructor
    B, J, S, K = 2, 17, 128, 10
    sm = torch.from_numpy(np.stack([synthetic.synth_scoremap(b, J, S, K) for b in range(B)])).to(DEV)
    backbone = torch.randn(B, 32, S, S, device=DEV).requires_grad_(True)
    conv = torch.nn.Conv2d(32, 128, 3, 1, 1).to(DEV)
    gcfg = pgmp_b200.config.bench_gc_config(k=K, graph_type="knn")
    ret = get_graph_constructor(gcfg, scoremaps=sm, tagmaps=None, features=ConvUpsampleFeatures(backbone, conv, (S, S)),
                                joints_gt=None, factor_list=None, masks=None, device=DEV, testing=True, heatmaps=None,
                                num_joints=J).construct_graph()
    x = ret[0]
    assert x.requires_grad
    x.square().sum().backward()
    assert backbone.grad is not None and conv.weight.grad is not None and conv.bias.grad is not None
    assert float(backbone.grad.abs().sum()) > 0 and float(conv.weight.grad.abs().sum()) > 0
    touched = (backbone.grad.abs().sum(1) > 0).sum().item()
    assert touched <= 9 * x.shape[0]                         # only the 3 x 3 neighbourhoods of the candidates (same size: one tap)
