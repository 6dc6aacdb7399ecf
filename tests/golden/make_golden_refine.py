"""Fixtures for the pose-assembly tail: the reference's own ``refine`` (src/Utils/Utils.py:1026-1104)
and ``adjust`` (:917-936), executed from their source lines (the module itself pulls matplotlib / tensorboard / the missing
native solver), on seeded synthetic inputs.

    python tests/golden/make_golden_refine.py        # needs /root/reference; the fixtures travel
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from cases import REFINE_CASES, refine_inputs  # noqa: E402


def load_reference():
    src = open("/root/reference/src/Utils/Utils.py").read().splitlines()
    ns = {"np": np}
    exec("\n".join(src[916:936]), ns)         # adjust
    exec("\n".join(src[1025:1104]), ns)       # refine
    return ns["refine"], ns["adjust"]


def main():
    refine, adjust = load_reference()
    for name in REFINE_CASES:
        sm, tags, kps = refine_inputs(name)
        out = {}
        for b in range(sm.shape[0]):
            r = refine(sm[b], tags[b], kps[b].copy())
            out[f"refined_{b}"] = r
            out[f"adjusted_{b}"] = adjust(r.copy(), sm[b])
            out[f"adjusted_only_{b}"] = adjust(kps[b].copy(), sm[b])
        np.savez_compressed(os.path.join(HERE, f"refine_{name}.npz"), **out)
        print(name, [out[f"refined_{b}"].shape for b in range(sm.shape[0])])


if __name__ == "__main__":
    main()
