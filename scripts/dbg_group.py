"""One GC + MPN + grouping pass on the bench workload (for ncu captures of group_kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pgmp_b200, pgmp_b200.synthetic as synthetic
from pgmp_b200.graph_constructor import get_graph_constructor
from pgmp_b200.Models.MessagePassingNetwork import get_mpn_model
from pgmp_b200.Utils import group_persons
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
J, K, S = 17, 30, 512
dev = "cuda:0"
sm = torch.from_numpy(np.stack([synthetic.synth_scoremap(b, J, S, K) for b in range(B)])).to(dev)
g = torch.Generator(device=dev).manual_seed(0)
feat = torch.randn(B, 128, S, S, device=dev, generator=g)
tags = torch.randn(B, J, S, S, device=dev, generator=g)
gcfg = pgmp_b200.config.bench_gc_config(k=K, graph_type="knn")
mcfg = pgmp_b200.config.flagship_mpn_config(J, B200_PRECISION="tc")
model = synthetic.synth_mpn_state_dict(get_mpn_model(mcfg), 1).eval().to(dev)
gc = get_graph_constructor(gcfg, scoremaps=sm, tagmaps=tags, features=feat, joints_gt=None, factor_list=None, masks=None,
                           device=dev, testing=True, heatmaps=None, num_joints=J)
ret = gc.construct_graph()
with torch.no_grad():
    pe, pn, pc, _ = model(ret[0], ret[1], ret[2], node_types=ret[7][:, 2])
for _ in range(2):
    res = group_persons(ret[7], pn[-1], ret[2], pe[-1], pc[-1], ret[12], J, node_threshold=0.1, detector_scores=ret[11],
                        nodes_per_image=gc.num_nodes_per_image, edges_per_image=gc.num_edges_per_image)
torch.cuda.synchronize()
print("ok", sum(0 if r is None else len(r[0]) for r in res))
