"""Node features at the candidate pixels without materialising the feature maps (numpy).
TEST INFRASTRUCTURE -- see oracle/__init__.py.

Restates what the reference computes between the backbone and the graph constructor
(SURVEY.md 8f, rank 1):

    feat = self.feature_gather(feat)                                  # Conv2d(Cin, 128, 3, 1, 1), PoseEstimation.py:64-66, 79, 341
    feat = interpolate(feat, size=(H, W), mode="bilinear", align_corners=False)   # PoseEstimation.py:442-450
    x = feat[:, y, x].T                                               # ConstructGraph.py:265, 269

Both steps are linear, so ``x[n] = sum over the 4 bilinear taps of weight * conv(feat)[:, tap]``; only the
candidate pixels are evaluated.  ``features_at_candidates_dense`` is the literal restatement (whole maps),
``features_at_candidates`` the candidate-only form the CUDA kernel follows.
"""

import numpy as np


def conv3x3(feat, weight, bias):
    """``nn.Conv2d(Cin, Cout, 3, stride 1, padding 1)`` on one image.  feat [Cin,h,w], weight [Cout,Cin,3,3]."""
    feat = np.asarray(feat, np.float32)
    cin, h, w = feat.shape
    pad = np.zeros((cin, h + 2, w + 2), np.float32)
    pad[:, 1:h + 1, 1:w + 1] = feat
    out = np.zeros((weight.shape[0], h, w), np.float32)
    for ky in range(3):
        for kx in range(3):
            out += np.einsum("oc,chw->ohw", weight[:, :, ky, kx].astype(np.float32), pad[:, ky:ky + h, kx:kx + w],
                             dtype=np.float32)
    return out + np.asarray(bias, np.float32)[:, None, None]


def bilinear_taps(dst, in_size, out_size):
    """Source indices and weights of ``interpolate(mode="bilinear", align_corners=False)`` along one axis
    (ATen ``area_pixel_compute_source_index``): src = max((dst + 0.5) * in/out - 0.5, 0)."""
    scale = np.float32(in_size) / np.float32(out_size)
    src = np.maximum((np.asarray(dst, np.float32) + np.float32(0.5)) * scale - np.float32(0.5), np.float32(0))
    i0 = np.minimum(np.floor(src).astype(np.int64), in_size - 1)
    i1 = np.minimum(i0 + 1, in_size - 1)
    l1 = (src - i0.astype(np.float32)).astype(np.float32)
    l0 = (np.float32(1) - l1).astype(np.float32)
    return i0, i1, l0, l1


def upsample_bilinear(maps, out_h, out_w):
    """[C,h,w] -> [C,out_h,out_w], align_corners=False."""
    maps = np.asarray(maps, np.float32)
    _, h, w = maps.shape
    y0, y1, ly0, ly1 = bilinear_taps(np.arange(out_h), h, out_h)
    x0, x1, lx0, lx1 = bilinear_taps(np.arange(out_w), w, out_w)
    top = maps[:, y0][:, :, x0] * lx0 + maps[:, y0][:, :, x1] * lx1
    bot = maps[:, y1][:, :, x0] * lx0 + maps[:, y1][:, :, x1] * lx1
    return (top * ly0[None, :, None] + bot * ly1[None, :, None]).astype(np.float32)


def features_at_candidates_dense(feat, weight, bias, joint_det, batch_index, out_h, out_w):
    """Literal restatement: conv on the whole map, upsample the whole map, gather."""
    n = joint_det.shape[0]
    x = np.zeros((n, weight.shape[0]), np.float32)
    for b in range(feat.shape[0]):
        sel = np.flatnonzero(batch_index == b)
        if sel.size == 0:
            continue
        up = upsample_bilinear(conv3x3(feat[b], weight, bias), out_h, out_w)
        x[sel] = up[:, joint_det[sel, 1], joint_det[sel, 0]].T
    return x


def features_at_candidates(feat, weight, bias, joint_det, batch_index, out_h, out_w):
    """Candidate-only form: the 3x3xCin input patch is interpolated (4 taps, zero padding applied per tap), then
    one [9 Cin] x [9 Cin, Cout] product per candidate."""
    feat = np.asarray(feat, np.float32)
    _, cin, h, w = feat.shape
    cout = weight.shape[0]
    wmat = np.asarray(weight, np.float32).transpose(2, 3, 1, 0).reshape(9 * cin, cout)   # [(ky,kx,ci), co]
    pad = np.zeros((feat.shape[0], cin, h + 2, w + 2), np.float32)
    pad[:, :, 1:h + 1, 1:w + 1] = feat
    y0, y1, ly0, ly1 = bilinear_taps(joint_det[:, 1], h, out_h)
    x0, x1, lx0, lx1 = bilinear_taps(joint_det[:, 0], w, out_w)
    x = np.zeros((joint_det.shape[0], cout), np.float32)
    for n in range(joint_det.shape[0]):
        b = int(batch_index[n])
        patch = np.zeros((3, 3, cin), np.float32)
        for (yy, wy) in ((y0[n], ly0[n]), (y1[n], ly1[n])):
            for (xx, wx) in ((x0[n], lx0[n]), (x1[n], lx1[n])):
                patch += np.float32(wy * wx) * pad[b, :, yy:yy + 3, xx:xx + 3].transpose(1, 2, 0)
        x[n] = patch.reshape(-1) @ wmat + np.asarray(bias, np.float32)
    return x
