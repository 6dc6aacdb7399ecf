"""BASELINE.json configs[2..3] sanity + timing sweep (development aid): w48-640 fully connected, CrowdPose-shaped."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pgmp_b200, pgmp_b200.synthetic as synthetic, pgmp_b200._native as nv
from pgmp_b200.graph_constructor import get_graph_constructor
from pgmp_b200.Models.MessagePassingNetwork import get_mpn_model
from pgmp_b200.Utils import group_persons

dev = "cuda:0"

def run(name, B, J, S, K, graph, prec="tc", steps=3):
    sm = torch.from_numpy(np.stack([synthetic.synth_scoremap(b, J, S, K, persons=8 if J == 17 else 20) for b in range(B)])).to(dev)
    g = torch.Generator(device=dev).manual_seed(0)
    feat = torch.randn(B, 128, S, S, device=dev, generator=g)
    tags = torch.randn(B, J, S, S, device=dev, generator=g)
    gcfg = pgmp_b200.config.bench_gc_config(k=K, graph_type=graph)
    over = {} if J == 17 else dict(NUM_JOINTS=J, EDGE_INPUT_DIM=J + 2)
    mcfg = pgmp_b200.config.flagship_mpn_config(J, B200_PRECISION=prec, **over)
    if J != 17:
        mcfg.CLASS.OUTPUT_SIZES = [64, 32, J]
    model = synthetic.synth_mpn_state_dict(get_mpn_model(mcfg), 1).eval().to(dev)

    def step():
        ret = get_graph_constructor(gcfg, scoremaps=sm, tagmaps=tags, features=feat, joints_gt=None, factor_list=None,
                                    masks=None, device=dev, testing=True, heatmaps=None, num_joints=J).construct_graph()
        with torch.no_grad():
            pe, pn, pc, _ = model(ret[0], ret[1], ret[2], node_types=ret[7][:, 2])
        return ret, pe, pn, pc

    for _ in range(2):
        ret, pe, pn, pc = step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        ret, pe, pn, pc = step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    N, E = ret[0].shape[0], ret[2].shape[1]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = group_persons(ret[7], pn[-1], ret[2], pe[-1], pc[-1], ret[12], J, node_threshold=0.1, detector_scores=ret[11])
    torch.cuda.synchronize()
    tg = time.perf_counter() - t0
    npers = [0 if r is None else len(r[0]) for r in res]
    print(f"{name}: B={B} N={N} E={E} gc+mpn {dt*1e3:.2f} ms/step {B/dt:.0f} img/s {E/dt/1e6:.1f} Medges/s | grouping {tg*1e3:.1f} ms, persons/img {np.mean(npers):.1f} | finite {bool(torch.isfinite(pe[-1]).all())}")

run("config2 512 knn", 32, 17, 512, 30, "knn")
run("config3 w48-640 fully", 8, 17, 640, 30, "fully")
run("config3 w48-640 knn", 8, 17, 640, 30, "knn")
run("config4 crowdpose knn", 8, 14, 512, 60, "knn")
run("config4 crowdpose fully", 4, 14, 512, 60, "fully")
run("config2 fp32 mode", 8, 17, 512, 30, "knn", prec="fp32")
