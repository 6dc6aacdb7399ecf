"""Fixtures for the last two steps of ``pred_to_ann``: the reference's own ``reverse_affine_map`` /
``reverse_affine_map_points`` (src/Utils/transformations.py, imported as a module: it needs cv2 and torchvision, both in
this container) and ``gen_ann_format*`` (src/Utils/eval.py:189-253, executed from their source lines: the module pulls
pycocotools) on seeded keypoints.

    python tests/golden/make_golden_transform.py        # needs /root/reference; the fixture travels
"""
import importlib.util
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# (image width, image height, input size, scaling type, min scale)
CASES = [(640, 480, 512, "short", 1.0), (427, 640, 512, "short", 1.0), (500, 375, 512, "short_with_resize", 1.0),
         (333, 500, 640, "short", 1.0), (640, 427, 512, "short", 0.5), (612, 612, 512, "short_with_resize", 2.0),
         (640, 480, 512, "long", 1.0), (480, 640, 512, "long_with_multiscale", 1.0)]
POINT_CASES = [(640, 480, "short", 1.0), (375, 500, "short_with_resize", 1.0), (640, 427, "short", 0.5)]


def keypoints(seed, persons=5, joints=17):
    rng = np.random.default_rng(seed)
    k = np.zeros((persons, joints, 3))
    k[:, :, :2] = rng.uniform(0, 320, (persons, joints, 2))
    k[:, :, 2] = rng.uniform(0, 1, (persons, joints)) * (rng.uniform(0, 1, (persons, joints)) > 0.3)
    k[-1, :, 2] = rng.uniform(0, 0.08, joints)          # a person without a visible joint
    return k


def main():
    spec = importlib.util.spec_from_file_location("ref_transformations", "/root/reference/src/Utils/transformations.py")
    T = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(T)
    src = open("/root/reference/src/Utils/eval.py").read().splitlines()
    ns = {"np": np}
    exec("\n".join(src[188:253]), ns)
    out = {"cases": [], "points": [], "ann": {}}
    for i, (w, h, inp, st, ms) in enumerate(CASES):
        k = keypoints(i)
        r = T.reverse_affine_map(k.copy(), (w, h), inp, scaling_type=st, min_scale=ms)
        out["cases"].append({"args": [w, h, inp, st, ms], "seed": i, "out": r.tolist()})
    for i, (w, h, st, ms) in enumerate(POINT_CASES):
        p = keypoints(100 + i)[0]
        r = T.reverse_affine_map_points(p.copy(), (w, h), scaling_type=st, min_scale=ms)
        out["points"].append({"args": [w, h, st, ms], "seed": 100 + i, "out": r.tolist()})
    k = keypoints(7)
    for name in ("gen_ann_format", "gen_ann_format_mean", "gen_ann_format_correct"):
        out["ann"][name] = ns[name](k, image_id=42)
    s, c, sc = T.get_multi_scale_size(480, 640, 512, 1.0, 1.0)
    out["multi_scale_size"] = {"size": list(s), "center": c.tolist(), "scale": sc.tolist()}
    json.dump(out, open(os.path.join(HERE, "transform.json"), "w"))
    print("wrote transform.json:", len(out["cases"]), "maps,", len(out["points"]), "point maps")


if __name__ == "__main__":
    main()
