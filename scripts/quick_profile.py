"""Per-kernel CUDA-event times of one GC + MPN pass (development aid)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pgmp_b200, pgmp_b200.synthetic as synthetic, pgmp_b200._native as nv
from pgmp_b200.graph_constructor import get_graph_constructor
from pgmp_b200.Models.MessagePassingNetwork import get_mpn_model

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
graph = sys.argv[2] if len(sys.argv) > 2 else "knn"
prec = sys.argv[3] if len(sys.argv) > 3 else "fp32"
dev = "cuda:0"
J, K, S = 17, 30, 512
sm = torch.from_numpy(np.stack([synthetic.synth_scoremap(b, J, S, K) for b in range(B)])).to(dev)
g = torch.Generator(device=dev).manual_seed(0)
feat = torch.randn(B, 128, S, S, device=dev, generator=g)
tags = torch.randn(B, J, S, S, device=dev, generator=g)
gcfg = pgmp_b200.config.bench_gc_config(k=K, graph_type=graph)
mcfg = pgmp_b200.config.flagship_mpn_config(J, B200_PRECISION=prec)
model = synthetic.synth_mpn_state_dict(get_mpn_model(mcfg), 1).eval().to(dev)

def step():
    ret = get_graph_constructor(gcfg, scoremaps=sm, tagmaps=tags, features=feat, joints_gt=None, factor_list=None,
                                masks=None, device=dev, testing=True, heatmaps=None, num_joints=J).construct_graph()
    with torch.no_grad():
        out = model(ret[0], ret[1], ret[2], node_types=ret[7][:, 2])
    return ret, out

for _ in range(3):
    ret, out = step()
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 5
for _ in range(n):
    ret, out = step()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n
N, E = ret[0].shape[0], ret[2].shape[1]
print(f"B={B} graph={graph} prec={prec} N={N} E={E}: {dt*1e3:.3f} ms/step  {B/dt:.1f} img/s  {E/dt/1e6:.2f} Medges/s")
nv.profile(True)
step()
prof = nv.profile_collect()
nv.profile(False)
tot = sum(v[1] for v in prof.values())
for k, (c, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:40s} x{c:3d} {ms:9.3f} ms  {100*ms/tot:5.1f}%")
print(f"  kernel total {tot:.3f} ms")
if len(sys.argv) > 4 and sys.argv[4] == "group":
    from pgmp_b200.Utils import group_persons
    args = (ret[7], out[1][-1], ret[2], out[0][-1], out[2][-1], ret[12], J)
    group_persons(*args, node_threshold=0.1, detector_scores=ret[11])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = group_persons(*args, node_threshold=0.1, detector_scores=ret[11])
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    nv.profile(True)
    group_persons(*args, node_threshold=0.1, detector_scores=ret[11])
    prof = nv.profile_collect()
    nv.profile(False)
    print(f"  grouping tail: wall {wall*1e3:.3f} ms; kernels {dict((k, round(v[1], 3)) for k, v in prof.items())}; "
          f"persons/img {np.mean([0 if r is None else len(r[0]) for r in res]):.1f}")
