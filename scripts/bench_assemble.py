"""Scoremap assembly: pgmp_gc_assemble_scoremaps against the reference's three torch operations
(interpolate + add + divide, hrnet.py:590-603) on the BASELINE shapes (32 x 17 x 512 x 512 from a 256 x 256 stage)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pgmp_b200  # noqa: E402,F401
import pgmp_b200._native as nv  # noqa: E402
from pgmp_b200.graph_constructor import hr_process_output  # noqa: E402

B, J, H = 32, 17, 512
s2 = torch.rand(B, J, H, H, device="cuda")
s1 = torch.rand(B, 2 * J, H // 2, H // 2, device="cuda")
for _ in range(3):
    hr_process_output(((s1, s2), None), "avg", J)
nv.profile(True)
for _ in range(10):
    hr_process_output(((s1, s2), None), "avg", J)
cnt, ms = nv.profile_collect()["assemble_kernel"]
nv.profile(False)
nbytes = 4 * (2 * B * J * H * H + B * J * H * H + B * 2 * J * (H // 2) ** 2)      # scores + tags written, stage2 + stage1 read
print("assemble_kernel: %.3f ms per launch, %.0f GB/s of algorithmic traffic (%.2f GB)" % (ms / cnt, nbytes / (ms / cnt) / 1e6, nbytes / 1e9))


def ref():
    up = torch.nn.functional.interpolate(s1, size=(H, H), mode="bilinear", align_corners=False)
    return (s2 + up[:, :J]) / 2, up[:, J:]


for _ in range(3):
    ref()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(10):
    ref()
ev1.record()
torch.cuda.synchronize()
print("stock torch operations: %.3f ms" % (ev0.elapsed_time(ev1) / 10))
