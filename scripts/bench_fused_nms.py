"""Scoremap assembly fused into the NMS loader (pgmp_gc_detect_fused) against the two-kernel path (assemble_kernel writes
the map, nms_candidates_kernel re-reads it) at the BASELINE size 32 x 17 x 512 x 512: per-kernel CUDA-event times."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pgmp_b200
import pgmp_b200._native as nv
import pgmp_b200.synthetic as synthetic
from pgmp_b200.graph_constructor import HeadStages, get_graph_constructor, hr_process_output

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
J, K, S = 17, 30, 512
dev = "cuda:0"
s2 = torch.from_numpy(np.stack([synthetic.synth_scoremap(b, J, S, K) for b in range(B)])).to(dev)
g = torch.Generator(device=dev).manual_seed(0)
heat = torch.nn.functional.avg_pool2d(s2, 2) + 0.01 * torch.rand(B, J, S // 2, S // 2, device=dev, generator=g)
s1 = torch.cat([heat, torch.randn(B, J, S // 2, S // 2, device=dev, generator=g)], 1).contiguous()
s1f, s2f = torch.flip(s1, [3]).contiguous(), torch.flip(s2, [3]).contiguous()
FLIP = [0, 2, 1, 4, 3, 6, 5, 8, 7, 10, 9, 12, 11, 14, 13, 16, 15]
feat = torch.zeros(B, 4, S, S, device=dev)
gcfg = pgmp_b200.config.bench_gc_config(k=K, graph_type="knn")


def gc(sm, tags):
    return get_graph_constructor(gcfg, scoremaps=sm, tagmaps=tags, features=feat, joints_gt=None, factor_list=None, masks=None,
                                 device=dev, testing=True, heatmaps=None, num_joints=J).construct_graph()


def two_kernels():
    score, _, tags = hr_process_output(((s1, s2), None), "avg", J)
    return gc(score, tags)


def fused(keep=False, flip=False):
    st = HeadStages((s1, s2), J, flipped=(s1f, s2f) if flip else None, flip_index=FLIP if flip else None, keep_scoremaps=keep)
    return gc(st, st)


out = {}
for name, fn in (("two_kernels", two_kernels), ("fused", fused), ("fused_keep_map", lambda: fused(True)),
                 ("fused_flip", lambda: fused(False, True))):
    for _ in range(3):
        fn()
    nv.profile(True)
    for _ in range(10):
        fn()
    prof = nv.profile_collect()
    nv.profile(False)
    ks = {k: v[1] / 10 for k, v in prof.items() if k.startswith("nms_candidates") or k.startswith("assemble")}
    out[name] = {"ms": sum(ks.values()), "kernels": ks}
    print(name, json.dumps(out[name]), flush=True)
alg = B * J * (S * S + 2 * (S // 2) ** 2 // 2) * 4   # stage 2 + the heatmap half of stage 1
print(json.dumps({"B": B, "bytes_read_fused": alg, "fused_GBps": alg / out["fused"]["ms"] / 1e6, **{k: v["ms"] for k, v in out.items()}}))
