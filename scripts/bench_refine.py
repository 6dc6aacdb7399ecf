"""Pose-assembly tail: pgmp_refine_persons on 32 images of 17 x 512 x 512 with 8 persons each, per-kernel
CUDA-event times, beside the numpy restatement of the reference's refine + adjust on the host for ONE image."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import oracle.refine as R  # noqa: E402  (measurement of the CPU side only)
import pgmp_b200  # noqa: E402,F401
import pgmp_b200._native as nv  # noqa: E402
from pgmp_b200.Utils import refine_persons  # noqa: E402

B, J, H, W, P = 32, 17, 512, 512, 8
rng = np.random.default_rng(0)
sm = torch.rand(B, J, H, W, device="cuda") * 0.3
tags = torch.randn(B, J, H, W, device="cuda") * 2
kps = []
for b in range(B):
    k = np.zeros((P, J, 3))
    k[:, :, 0], k[:, :, 1] = rng.integers(0, W, (P, J)), rng.integers(0, H, (P, J))
    k[:, :, 2] = np.where(rng.uniform(size=(P, J)) > 0.3, rng.uniform(0.1, 1, (P, J)), 0.0)
    k[:, 0, 2] = 0.5
    kps.append(k)
for _ in range(2):
    out = refine_persons(sm, tags, kps)
nv.profile(True)
out = refine_persons(sm, tags, kps)
prof = nv.profile_collect()
nv.profile(False)
tot = sum(v[1] for v in prof.values())
for name, (cnt, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    print("%-24s %8.3f ms" % (name, ms))
print("device total %.3f ms for %d images x %d persons (%.0f images/s); maps read once per 8 persons: %.2f GB" %
      (tot, B, P, B / tot * 1e3, 2 * B * J * H * W * 4 / 1e9))
s0, t0 = sm[0].cpu().numpy(), tags[0].cpu().numpy()
t = time.perf_counter()
want = R.adjust(R.refine(s0, t0, kps[0]), s0)
dt = time.perf_counter() - t
assert np.array_equal(out[0], want)
print("numpy refine + adjust on the host, one image: %.1f ms (%.1f images/s)" % (dt * 1e3, 1 / dt))
