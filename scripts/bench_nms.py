"""NMS kernel alone on the BASELINE workload (32 x 17 x 512 x 512): CUDA-event time per launch for several strip
heights (PGMP_NMS_STRIP_ROWS), achieved algorithmic GB/s.  Development aid; bench.py reports the default."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pgmp_b200
import pgmp_b200._native as nv
import pgmp_b200.synthetic as synthetic
from pgmp_b200.graph_constructor import get_graph_constructor

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
J, K, S = 17, 30, 512
dev = "cuda:0"
sm = torch.from_numpy(np.stack([synthetic.synth_scoremap(b, J, S, K) for b in range(B)])).to(dev)
feat = torch.zeros(B, 4, S, S, device=dev)
gcfg = pgmp_b200.config.bench_gc_config(k=K, graph_type="knn")
out = {}
for rows in [0, "notail", 128, 256, 512]:
    os.environ.pop("PGMP_NMS_STRIP_ROWS", None)
    os.environ.pop("PGMP_NMS_TAIL", None)
    if rows == "notail":          # whole maps only: 544 CTAs on 444 slots, the last wave is partial
        os.environ["PGMP_NMS_TAIL"] = "0"
    elif rows:
        os.environ["PGMP_NMS_STRIP_ROWS"] = str(rows)

    def step():
        return get_graph_constructor(gcfg, scoremaps=sm, tagmaps=sm, features=feat, joints_gt=None, factor_list=None, masks=None,
                                     device=dev, testing=True, heatmaps=None, num_joints=J).construct_graph()
    for _ in range(3):
        step()
    nv.profile(True)
    for _ in range(10):
        step()
    prof = nv.profile_collect()
    nv.profile(False)
    name = next(k for k in prof if k.startswith("nms_candidates"))
    ms = prof[name][1] / prof[name][0]
    out[rows or "default"] = {"ms": ms, "GBps": B * J * S * S * 4 / ms / 1e6,
                              "select_ms": prof["select_detections_kernel"][1] / prof["select_detections_kernel"][0]}
    print(rows or "default", out[rows or "default"], flush=True)
print(json.dumps(out))
