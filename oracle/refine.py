"""Pose-assembly tail after the grouping (numpy) -- TEST INFRASTRUCTURE, see oracle/__init__.py.

Restates ``refine`` (src/Utils/Utils.py:1026-1104: joints a person is missing are looked up in the heatmaps at the
position whose tag is closest to the person's mean tag) and ``adjust`` (:917-936: quarter-pixel offsets towards the
higher neighbour).  Pinned by tests/test_refine.py to the reference's own functions (tests/golden/refine_*.npz,
written by tests/golden/make_golden_refine.py).  The float32 details that decide the arg-max are kept: the mean tag is
numpy's float32 ``mean`` (pairwise over a contiguous axis when the tag dimension is 1, row by row otherwise), the
distance ``sqrt(sum((tag - mean)^2))`` is rounded half-to-even, ties go to the first pixel in row-major order."""
import numpy as np

f32 = np.float32


def mean_tag(vals):
    """np.mean(vals, axis=0) for float32 [n, T]."""
    return np.mean(np.asarray(vals, f32), axis=0)


def refine(scoremaps, tag, keypoints):
    """scoremaps [J,H,W] f32, tag [J,H,W] or [J,H,W,T] f32, keypoints [P,J,3] f64 (x, y, score) -> refined copy."""
    kp = np.array(keypoints, dtype=np.float64, copy=True)
    tag = np.asarray(tag, f32)
    if tag.ndim == 3:
        tag = tag[..., None]
    J, H, W = scoremaps.shape
    for p in range(kp.shape[0]):
        det = kp[p, :, 2] > 0
        xi, yi = kp[p, :, 0].astype(np.int32), kp[p, :, 1].astype(np.int32)       # Utils.py:1045
        prev = mean_tag(tag[np.arange(J)[det], yi[det], xi[det]])                   # :1062
        dist = np.sqrt(((tag - prev[None, None, None, :]) ** 2).sum(axis=3, dtype=f32)).astype(f32)   # :1070
        score = (scoremaps - np.round(dist)).astype(f32).reshape(J, -1)             # :1071
        flat = np.argmax(score, axis=1)                                             # :1074, first maximum
        y, x = np.unravel_index(flat, (H, W))
        val = scoremaps[np.arange(J), y, x]
        right = scoremaps[np.arange(J), y, np.minimum(x + 1, W - 1)] > scoremaps[np.arange(J), y, np.maximum(x - 1, 0)]
        down = scoremaps[np.arange(J), np.minimum(y + 1, H - 1), x] > scoremaps[np.arange(J), np.maximum(y - 1, 0), x]
        fx = x + 0.5 + np.where(right, 0.25, -0.25)                                 # :1080-1091
        fy = y + 0.5 + np.where(down, 0.25, -0.25)
        add = (val > 0) & (kp[p, :, 2] == 0)                                        # :1098-1101
        kp[p, add, 0], kp[p, add, 1], kp[p, add, 2] = fx[add], fy[add], 0.001
    return kp


def adjust(keypoints, scoremaps):
    """Utils.py:917-936 (``ans`` holds (x, y, score); the reference's local names y / x are swapped, the arithmetic is not)."""
    kp = np.array(keypoints, dtype=np.float64, copy=True)
    J, H, W = scoremaps.shape
    for p in range(kp.shape[0]):
        for j in range(J):
            if kp[p, j, 2] > 0:
                cx, cy = kp[p, j, 0], kp[p, j, 1]
                col, row = int(cx), int(cy)
                m = scoremaps[j]
                cx += 0.25 if m[row, min(col + 1, W - 1)] > m[row, max(col - 1, 0)] else -0.25
                cy += 0.25 if m[min(row + 1, H - 1), col] > m[max(0, row - 1), col] else -0.25
                kp[p, j, 0], kp[p, j, 1] = cx + 0.5, cy + 0.5
    return kp
