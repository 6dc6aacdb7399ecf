"""Development aid: host-side timing of the training-time label construction (8 images of 256 x 256)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pgmp_b200, pgmp_b200.synthetic as synthetic
from pgmp_b200.graph_constructor import get_graph_constructor, labels as L

dev = "cuda:0"
J, K, B, S = 17, 30, 8, 256
data = synthetic.synth_batch(B, J, S, K, persons=8)
t = {k: torch.from_numpy(v).to(dev) for k, v in data.items()}
gts, facs = zip(*[synthetic.synth_joints_gt(b, J, S, K, persons=8) for b in range(B)])
joints_gt, factors = torch.from_numpy(np.stack(gts)).to(dev), torch.from_numpy(np.stack(facs)).to(dev)
gcfg = pgmp_b200.config.bench_gc_config(k=K, graph_type="knn", EDGE_LABEL_METHOD=6, MATCHING_RADIUS=0.5)
print("torch threads", torch.get_num_threads(), "cpus", os.cpu_count())
def run(joints):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ret = get_graph_constructor(gcfg, scoremaps=t["scoremaps"], tagmaps=t["tagmaps"], features=t["features"], joints_gt=joints,
                                factor_list=factors if joints is not None else None, masks=None, device=dev, testing=False, heatmaps=None, num_joints=J).construct_graph()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3
for _ in range(3): run(joints_gt)
print("construct_graph with labels ms", [round(run(joints_gt), 2) for _ in range(6)])
print("construct_graph without    ms", [round(run(None), 2) for _ in range(6)])
# parts
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5): run(joints_gt)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
