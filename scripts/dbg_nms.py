import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pgmp_b200, pgmp_b200._native as nv, pgmp_b200.synthetic as synthetic
from pgmp_b200.graph_constructor import get_graph_constructor
J, K, S = 17, 30, 512
dev = "cuda:0"
once = "--once" in sys.argv
Bs = [int(a) for a in sys.argv[1:] if a != "--once"] or [8, 16, 32]
sm_all = torch.from_numpy(np.stack([synthetic.synth_scoremap(b, J, S, K) for b in range(max(Bs))])).to(dev)
gcfg = pgmp_b200.config.bench_gc_config(k=K, graph_type="knn")
for B in Bs:
    for rows in ([0, 128, 256, 512] if not once else [int(os.environ.get("PGMP_NMS_STRIP_ROWS", "0"))]):
        if rows: os.environ["PGMP_NMS_STRIP_ROWS"] = str(rows)
        else: os.environ.pop("PGMP_NMS_STRIP_ROWS", None)
        sm = sm_all[:B].contiguous()
        feat = torch.zeros(B, 4, S, S, device=dev)
        try:
            ret = get_graph_constructor(gcfg, scoremaps=sm, tagmaps=sm, features=feat, joints_gt=None, factor_list=None, masks=None,
                                        device=dev, testing=True, heatmaps=None, num_joints=J).construct_graph()
            torch.cuda.synchronize()
            print("B", B, "rows", rows, "ok N", ret[0].shape[0], flush=True)
        except Exception as e:
            print("B", B, "rows", rows, "FAILED", str(e)[:200], flush=True)
            sys.exit(1)
