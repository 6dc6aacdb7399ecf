"""Host steps at the end of ``pred_to_ann`` (Utils.py:1478-1489) against fixtures from the reference's own functions
(tests/golden/make_golden_transform.py): keypoints back to source-image coordinates, annotation records."""
import importlib.util
import json
import os

import numpy as np
import pytest

import pgmp_b200
from pgmp_b200.Utils import transformations as T

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "transform.json")))
spec = importlib.util.spec_from_file_location("make_golden_transform", os.path.join(HERE, "golden", "make_golden_transform.py"))
gen = importlib.util.module_from_spec(spec)
spec.loader.exec_module(gen)


@pytest.mark.parametrize("case", GOLD["cases"], ids=lambda c: "-".join(map(str, c["args"])))
def test_reverse_affine_map_matches_the_reference(case):
    w, h, inp, st, ms = case["args"]
    k = gen.keypoints(case["seed"])
    got = T.reverse_affine_map(k.copy(), (w, h), inp, scaling_type=st, min_scale=ms)
    want = np.array(case["out"])
    assert got.shape == want.shape
    np.testing.assert_array_equal(got[:, :, 2], want[:, :, 2])                 # scores untouched
    np.testing.assert_allclose(got[:, :, :2], want[:, :, :2], rtol=1e-9, atol=1e-9)   # float64 solve vs OpenCV's


@pytest.mark.parametrize("case", GOLD["points"], ids=lambda c: "-".join(map(str, c["args"])))
def test_reverse_affine_map_points_matches_the_reference(case):
    w, h, st, ms = case["args"]
    p = gen.keypoints(case["seed"])[0]
    got = T.reverse_affine_map_points(p.copy(), (w, h), scaling_type=st, min_scale=ms)
    np.testing.assert_allclose(got, np.array(case["out"]), rtol=1e-9, atol=1e-9)


def test_annotation_records_are_identical():
    k = gen.keypoints(7)
    for name, want in GOLD["ann"].items():
        got = getattr(T, name)(k, image_id=42)
        assert got == want, name                                               # plain Python floats: exact
    s, c, sc = T.get_multi_scale_size(480, 640, 512, 1.0, 1.0)
    assert list(s) == GOLD["multi_scale_size"]["size"] and c.tolist() == GOLD["multi_scale_size"]["center"]
    assert sc.tolist() == GOLD["multi_scale_size"]["scale"]


def test_persons_to_ann_chain_and_errors():
    k = gen.keypoints(3)
    ann = T.persons_to_ann(k, (640, 480), 512, 9, "short", 1.0, "mean")
    ref = T.gen_ann_format_mean(T.reverse_affine_map(k.copy(), (640, 480), 512, "short", 1.0), 9)
    assert ann == ref and len(ann) == len(k) and ann[0]["image_id"] == 9
    assert T.persons_to_ann(None, (640, 480), 512, 9, "short") is None
    with pytest.raises(NotImplementedError):
        T.reverse_affine_map(k.copy(), (640, 480), 512, "diagonal")
    with pytest.raises(NotImplementedError):
        T.persons_to_ann(k, (640, 480), 512, 9, "short", scoring_method="median")
    with pytest.raises(AssertionError):
        T.reverse_affine_map(k.copy(), (640, 480), 640, "long")
    # the inverse of the forward map: a keypoint at the output centre lands on the image centre
    size, center, scale = T.get_multi_scale_size(480, 640, 512, 1.0, 1.0)
    mid = np.array([[[size[0] / 4.0, size[1] / 4.0, 1.0]]])
    back = T.reverse_affine_map(mid.copy(), (640, 480), 512, "short")
    np.testing.assert_allclose(back[0, 0, :2], center, atol=1e-6)
