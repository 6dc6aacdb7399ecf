"""Development aid: clock64 timeline of the step kernel's phases (library built with PGMP_NVCC_EXTRA=-DPGMP_TIMELINE)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pgmp_b200, pgmp_b200.synthetic as synthetic, pgmp_b200._native as nv
from pgmp_b200.graph_constructor import get_graph_constructor
from pgmp_b200.Models.MessagePassingNetwork import get_mpn_model

B, J, K, S = 32, 17, 30, 512
dev = "cuda:0"
sm = torch.from_numpy(np.stack([synthetic.synth_scoremap(b, J, S, K) for b in range(B)])).to(dev)
g = torch.Generator(device=dev).manual_seed(0)
feat = torch.randn(B, 128, S, S, device=dev, generator=g)
tags = torch.randn(B, J, S, S, device=dev, generator=g)
gcfg = pgmp_b200.config.bench_gc_config(k=K, graph_type="knn")
mcfg = pgmp_b200.config.flagship_mpn_config(J, B200_PRECISION="tc")
model = synthetic.synth_mpn_state_dict(get_mpn_model(mcfg), 1).eval().to(dev)
for _ in range(3):
    ret = get_graph_constructor(gcfg, scoremaps=sm, tagmaps=tags, features=feat, joints_gt=None, factor_list=None,
                                masks=None, device=dev, testing=True, heatmaps=None, num_joints=J).construct_graph()
    with torch.no_grad():
        out = model(ret[0], ret[1], ret[2], node_types=ret[7][:, 2])
torch.cuda.synchronize()
buf = np.zeros((4, 32, 32), dtype=np.int64)
fn = nv.lib().pgmp_debug_step_timeline
fn.restype, fn.argtypes = C.c_int, [C.c_void_p]
assert fn(buf.ctypes.data) == 0
names = ["start", "PQ issued", "PQ summed", "C landed", "MMA1 done", "epi1 done", "bar", "R issued", "MMA2 done", "epi2 done",
         "bar", "scan done", "MMA3 done", "msg stored", "bar", "reduced"]
for slot in range(4):
    t = buf[slot]
    print(f"--- CTA {'3' if slot < 2 else '100'} tile group {slot & 1}")
    order = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 16, 17, 18, 19, 20, 11, 12, 13, 14, 21, 15]
    lab = ["start", "PQissue", "PQsum", "Cwait", "MMA1w", "epi1", "bar", "MMA2i+R", "MMA2w", "epi2", "bar", "MMA3iss", "scan-a", "scanbar",
           "scan-b", "ctl/cw", "atomics", "MMA3w", "msg", "bar", "g-issue", "reduce", "endbar"]
    rows = []
    for i in range(4, 20):
        pts = [t[i][k] for k in order] + [t[i + 1][0]]
        rows.append(np.diff(pts))
    rows = np.stack(rows)
    print("      " + " ".join(f"{n[:7]:>7s}" for n in lab[1:]))
    for i in range(4):
        print(f"t{i+4:3d}  " + " ".join(f"{int(x):7d}" for x in rows[i]), " total", rows[i].sum())
    print("mean  " + " ".join(f"{x:7.0f}" for x in rows.mean(0)), " total", rows.sum(1).mean())
# offset of the two groups of a CTA
for c in (0, 2):
    print("group offset (start of tile i of group 1 - group 0):", [int(buf[c + 1][i][0] - buf[c][i][0]) for i in range(6, 12)])
