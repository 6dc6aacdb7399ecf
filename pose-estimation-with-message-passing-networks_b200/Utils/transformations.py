"""Keypoints from network coordinates back to the source image, and the annotation records: the last two steps of
``pred_to_ann`` (``src/Utils/Utils.py:1478-1489``).

Host code on a few dozen floats per image (float64, like the reference).  ``reverse_affine_map`` restates
``src/Utils/transformations.py:7-82``; the reference builds the inverse map with ``cv2.getAffineTransform`` from three
point pairs -- here the same three pairs (rounded to float32 as the reference rounds them) are solved in float64 with
numpy, no OpenCV dependency; results agree to ~1e-12 relative (``tests/test_transformations.py``, fixtures from the
reference's own functions).  ``gen_ann_format`` / ``_mean`` / ``_correct`` restate ``src/Utils/eval.py:189-253``.
"""
import numpy as np

__all__ = ["get_multi_scale_size", "get_transform", "get_affine_transform", "kpt_affine", "reverse_affine_map",
           "reverse_affine_map_points", "gen_ann_format", "gen_ann_format_mean", "gen_ann_format_correct", "persons_to_ann"]


def get_multi_scale_size(img_h, img_w, input_size, current_scale, min_scale):
    """Network input size, centre and scale of an image whose short side is resized to ``input_size`` and whose long side
    is padded up to a multiple of 64 (``transformations.py:216-237``)."""
    h, w = img_h, img_w
    center = np.array([int(w / 2.0 + 0.5), int(h / 2.0 + 0.5)])
    base = int((min_scale * input_size + 63) // 64 * 64)
    factor = current_scale / min_scale
    if w < h:
        w_res = int(base * factor)
        h_res = int(int((base / w * h + 63) // 64 * 64) * factor)
        scale = (w / 200.0, h_res / w_res * w / 200.0)
    else:
        h_res = int(base * factor)
        w_res = int(int((base / h * w + 63) // 64 * 64) * factor)
        scale = (w_res / h_res * h / 200.0, h / 200.0)
    return (w_res, h_res), center, np.array(scale)


def get_transform(center, scale, res):
    """The 3 x 3 crop matrix of the hourglass pipeline without rotation (``transformations.py:142-151``)."""
    h = 200 * np.asarray(scale, dtype=np.float64)
    t = np.zeros((3, 3))
    t[0, 0] = float(res[1]) / h[1]
    t[1, 1] = float(res[0]) / h[0]
    t[0, 2] = res[1] * (-float(center[0]) / h[0] + .5)
    t[1, 2] = res[0] * (-float(center[1]) / h[1] + .5)
    t[2, 2] = 1
    return t


def _affine_from_points(src, dst):
    """The 2 x 3 matrix M with M [x y 1]^T = dst for three point pairs (what ``cv2.getAffineTransform`` solves)."""
    a = np.concatenate([np.asarray(src, np.float64), np.ones((3, 1))], axis=1)
    return np.linalg.solve(a, np.asarray(dst, np.float64)).T


def get_affine_transform(center, scale, output_size, inv=False):
    """``transformations.py:170-213``: centre, a point half the source width above it and the point at a right angle to
    both, mapped onto the corresponding points of the output; the pairs are float32 like the reference's."""
    scale = np.asarray(scale, dtype=np.float64)
    if scale.ndim == 0:
        scale = np.array([scale, scale])
    src_w = scale[0] * 200.0
    dst_w, dst_h = output_size[0], output_size[1]
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0] = center
    src[1] = np.asarray(center) + np.array([0.0, src_w * -0.5])
    dst[0] = [dst_w * 0.5, dst_h * 0.5]
    dst[1] = np.array([dst_w * 0.5, dst_h * 0.5]) + np.array([0, dst_w * -0.5], np.float32)
    for p in (src, dst):                       # third point: the second one turned by 90 degrees about the first
        d = p[0] - p[1]
        p[2] = p[1] + np.array([-d[1], d[0]], dtype=np.float32)
    return _affine_from_points(dst, src) if inv else _affine_from_points(src, dst)


def kpt_affine(kpt, mat):
    """``transformations.py:129-133``."""
    kpt = np.array(kpt)
    flat = kpt.reshape(-1, 2)
    return np.dot(np.concatenate((flat, flat[:, 0:1] * 0 + 1), axis=1), mat.T).reshape(kpt.shape)


def _inverse_map(img_size_orig, input_size, scaling_type, min_scale, long_res):
    """(2 x 3 inverse matrix, factor applied to the keypoints first) of a scaling type."""
    if scaling_type == "short":                # half-resolution output of the network
        size, center, scale = get_multi_scale_size(img_size_orig[1], img_size_orig[0], input_size, 1., min_scale)
        return get_affine_transform(center, scale, (int(size[0] / 2), int(size[1] / 2)), inv=True), 1
    if scaling_type == "short_with_resize":    # maps resized to the input resolution
        size, center, scale = get_multi_scale_size(img_size_orig[1], img_size_orig[0], input_size, 1., min_scale)
        return get_affine_transform(center, scale, (int(size[0]), int(size[1])), inv=True), 1
    if scaling_type in ("long", "long_with_multiscale"):
        if input_size != 512:
            raise AssertionError("scaling_type %r is defined for input_size 512" % scaling_type)
        w, h = img_size_orig[0], img_size_orig[1]
        s = max(h, w) / 200
        mat = get_transform(np.array((w / 2, h / 2)), np.array([s, s]), long_res[scaling_type])
        return np.linalg.pinv(mat)[:2], 4
    if scaling_type == "short_mine":
        size, center, scale = get_multi_scale_size(img_size_orig[1], img_size_orig[0], 512, 1., 1.)
        mat = get_transform(center, scale, (int(size[0] / 2), int(size[1] / 2)))
        return np.linalg.inv(mat)[:2], 1
    raise NotImplementedError(scaling_type)


def reverse_affine_map(keypoints, img_size_orig, input_size, scaling_type, min_scale=1.0):
    """``[P, J, 3]`` keypoints (x, y, score) in network coordinates -> source-image coordinates, in place like the
    reference (``transformations.py:7-82``).  ``img_size_orig`` = (width, height)."""
    mat, pre = _inverse_map(img_size_orig, input_size, scaling_type, min_scale,
                            {"long": (512, 512), "long_with_multiscale": (1024, 1024)})
    keypoints[:, :, :2] = kpt_affine(keypoints[:, :, :2] * pre if pre != 1 else keypoints[:, :, :2], mat)
    return keypoints


def reverse_affine_map_points(points, img_size_orig, scaling_type, min_scale=1.0):
    """The same for an ``[N, 3]`` array of detections (``transformations.py:85-126``; input size 512, the ``long`` type
    at the 128-pixel output resolution and without the factor 4)."""
    if scaling_type == "long_with_multiscale":
        raise NotImplementedError(scaling_type)
    mat, _ = _inverse_map(img_size_orig, 512, scaling_type, min_scale, {"long": (128, 128)})
    points[:, :2] = kpt_affine(points[:, :2], mat)
    return points


def _ann(pred, image_id, mean_part, sum_part):
    out = []
    for person in pred:
        score = 0.0
        if mean_part:
            seen = person[:, 2] > 0.09
            score = float(person[seen, 2].mean()) if seen.sum() > 0 else 0.0
        kps = []
        for j in range(len(person)):
            kps += [float(person[j, 0]), float(person[j, 1]), float(person[j, 2])]
            if sum_part:
                score += float(person[j, 2])
        out.append({"image_id": int(image_id), "category_id": 1, "keypoints": kps, "score": score})
    return out


def gen_ann_format(pred, image_id=0):
    """COCO keypoint records, score = mean of the visible joint scores + sum of all joint scores (``eval.py:189-211``)."""
    return _ann(pred, image_id, True, True)


def gen_ann_format_correct(pred, image_id=0):
    """Score = sum of the joint scores (``eval.py:213-231``)."""
    return _ann(pred, image_id, False, True)


def gen_ann_format_mean(pred, image_id=0):
    """Score = mean of the visible joint scores (``eval.py:233-253``)."""
    return _ann(pred, image_id, True, False)


def persons_to_ann(persons, img_shape, input_size, img_id, scaling_type, min_scale=1.0, scoring_method="default"):
    """The end of ``pred_to_ann`` (``Utils.py:1478-1489``) for one image: ``persons`` is the ``[P, J, 3]`` array
    ``persons_from_groups`` returns (or ``None`` -> ``None``, the reference returns no annotation)."""
    if persons is None:
        return None
    orig = reverse_affine_map(np.array(persons, dtype=np.float64, copy=True), img_shape, input_size,
                              scaling_type=scaling_type, min_scale=min_scale)
    fn = {"default": gen_ann_format, "mean": gen_ann_format_mean, "correct": gen_ann_format_correct}.get(scoring_method)
    if fn is None:
        raise NotImplementedError(scoring_method)
    return fn(orig, img_id)
