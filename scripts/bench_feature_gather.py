"""Timing of the lazy feature path against the reference's materialised path on one GPU."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pgmp_b200, pgmp_b200.synthetic as synthetic, pgmp_b200._native as nv
from pgmp_b200.graph_constructor import ConvUpsampleFeatures, get_graph_constructor

dev = "cuda:0"
B, J, S, K = 32, 17, 512, 30
sm = torch.from_numpy(np.stack([synthetic.synth_scoremap(b, J, S, K) for b in range(B)])).to(dev)
g = torch.Generator(device=dev).manual_seed(0)
backbone = torch.randn(B, 32, S // 2, S // 2, device=dev, generator=g)
tags = torch.randn(B, J, S, S, device=dev, generator=g)
conv = torch.nn.Conv2d(32, 128, 3, 1, 1).to(dev)
gcfg = pgmp_b200.config.bench_gc_config(k=K, graph_type="knn")


def gc(features):
    return get_graph_constructor(gcfg, scoremaps=sm, tagmaps=tags, features=features, joints_gt=None, factor_list=None,
                                 masks=None, device=dev, testing=True, heatmaps=None, num_joints=J).construct_graph()


def materialised():
    with torch.no_grad():
        full = torch.nn.functional.interpolate(conv(backbone), size=(S, S), mode="bilinear", align_corners=False)
    return gc(full)


def lazy():
    return gc(ConvUpsampleFeatures(backbone, conv, (S, S)))


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


t_mat, t_lazy = timeit(materialised), timeit(lazy)
nv.profile(True)
lazy()
prof = nv.profile_collect()
nv.profile(False)
print(f"B={B} w32 shapes (32 x 256^2 backbone map -> 128 x 512^2): conv + interpolate + construct_graph {t_mat:.3f} ms, "
      f"lazy construct_graph {t_lazy:.3f} ms; gather_conv_kernel {[v[1] for k, v in prof.items() if k.startswith('gather_conv_kernel')][0]:.3f} ms "
      f"(N = {B * J * K} candidates)")
