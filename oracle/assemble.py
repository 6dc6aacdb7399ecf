"""Scoremap assembly in front of the NMS (numpy, float32) -- TEST INFRASTRUCTURE, see oracle/__init__.py.

Restates ``hr_process_output`` (src/Models/HigherHRNet/hrnet.py:587-611): bilinear up-sampling of the half-resolution
stage with ``align_corners=False`` exactly as ATen computes it (``area_pixel_compute_source_index``: ``src = scale *
(dst + 0.5) - 0.5`` clamped at 0 with ``scale = in / out``; weights ``(1 - l, l)``; every product and sum rounded to
float32 separately), the average with the full-resolution stage, and the tag maps.  Pinned to torch's own
``interpolate`` on the CPU by ``tests/test_assemble.py`` (agreement to 1 ulp-level tolerance: ATen's vectorised CPU
kernel may contract a multiply-add)."""
import numpy as np

f32 = np.float32


def _source(scale, n_out, n_in):
    dst = np.arange(n_out, dtype=f32)
    src = (f32(scale) * (dst + f32(0.5))).astype(f32) - f32(0.5)
    src = np.maximum(src, f32(0.0)).astype(f32)
    i0 = np.minimum(src.astype(np.int64), n_in - 1)
    i1 = i0 + (i0 < n_in - 1)
    l1 = (src - i0.astype(f32)).astype(f32)
    l0 = (f32(1.0) - l1).astype(f32)
    return i0, i1, l0, l1


def upsample_bilinear(x, H, W):
    """``torch.nn.functional.interpolate(x, size=(H, W), mode='bilinear', align_corners=False)`` for [B, C, h, w] float32."""
    x = np.asarray(x, f32)
    h, w = x.shape[2], x.shape[3]
    y0, y1, wy0, wy1 = _source(f32(h) / f32(H), H, h)
    x0, x1, wx0, wx1 = _source(f32(w) / f32(W), W, w)
    top = (wx0 * x[:, :, y0][:, :, :, x0]).astype(f32) + (wx1 * x[:, :, y0][:, :, :, x1]).astype(f32)
    bot = (wx0 * x[:, :, y1][:, :, :, x0]).astype(f32) + (wx1 * x[:, :, y1][:, :, :, x1]).astype(f32)
    return ((wy0[:, None] * top.astype(f32)).astype(f32) + (wy1[:, None] * bot.astype(f32)).astype(f32)).astype(f32)


def hr_process_output(s1, s2, num_joints, mode="avg"):
    """-> (scoremaps [B, J, H, W], tags [B, C1 - J, H, W]); hrnet.py:590-608."""
    if mode == "large":
        return np.asarray(s2, f32), np.asarray(s1, f32)[:, num_joints:]
    up = upsample_bilinear(s1, s2.shape[2], s2.shape[3])
    tags = up[:, num_joints:]
    if mode == "avg":
        return ((np.asarray(s2, f32) + up[:, :num_joints]).astype(f32) * f32(0.5)).astype(f32), tags
    if mode == "small":
        return up[:, :num_joints], tags
    raise NotImplementedError(mode)


def flip_average(score, score_flipped, flip_index):
    """Flip-test average of one scale (PoseEstimation.py:377-402 + multi_scales_testing.py:162): the flipped image's
    assembled heatmaps are mirrored back along x (``torch.flip(output, [3])``), their joint channels permuted
    (``[:, flip_index]``) and averaged with the plain ones, ``(heatmaps[0] + heatmaps[1]) / 2.0``."""
    back = np.asarray(score_flipped, f32)[:, :, :, ::-1][:, list(flip_index)]
    return ((np.asarray(score, f32) + back).astype(f32) * f32(0.5)).astype(f32)
