import sys, os, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import pgmp_b200, pgmp_b200.synthetic as synthetic
from pgmp_b200.graph_constructor import get_graph_constructor
from pgmp_b200.Models.MessagePassingNetwork import get_mpn_model
dev="cuda:0"; B,J,S,K=32,17,512,30
sm = torch.from_numpy(np.stack([synthetic.synth_scoremap(b, J, S, K) for b in range(B)])).to(dev)
g = torch.Generator(device=dev).manual_seed(0)
feat = torch.randn(B,128,S,S,device=dev,generator=g); tags = torch.randn(B,J,S,S,device=dev,generator=g)
gcfg = pgmp_b200.config.bench_gc_config(k=K, graph_type="knn")
mcfg = pgmp_b200.config.flagship_mpn_config(J, B200_PRECISION="tc")
model = synthetic.synth_mpn_state_dict(get_mpn_model(mcfg), 1).eval().to(dev)
def step():
    t0=time.perf_counter()
    gc = get_graph_constructor(gcfg, scoremaps=sm, tagmaps=tags, features=feat, joints_gt=None, factor_list=None, masks=None, device=dev, testing=True, heatmaps=None, num_joints=J)
    ret = gc.construct_graph()
    t1=time.perf_counter()
    with torch.no_grad():
        out = model(ret[0],ret[1],ret[2],node_types=ret[7][:,2])
    t2=time.perf_counter()
    return t1-t0, t2-t1
for _ in range(3): step()
torch.cuda.synchronize()
a=[];b=[]
t0=time.perf_counter()
for _ in range(20):
    x,y=step(); a.append(x); b.append(y)
torch.cuda.synchronize()
tot=(time.perf_counter()-t0)/20
print(f"wall/step {tot*1e3:.3f} ms; host in construct_graph (incl. its sync) {np.mean(a)*1e3:.3f} ms; host in mpn forward {np.mean(b)*1e3:.3f} ms")

# where the host time goes (cProfile over 20 pipelined batches, cumulative top entries)
import cProfile, pstats, io
from pgmp_b200.pipeline import GroupingPipeline
pipe = GroupingPipeline(gcfg, model, J, dev)
batch = {"scoremaps": sm, "tagmaps": tags, "features": feat}
for _ in pipe.run(batch for _ in range(5)):
    pass
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in pipe.run(batch for _ in range(40)):
    pass
t_host = (time.perf_counter() - t0) / 40
torch.cuda.synchronize()
t_all = (time.perf_counter() - t0) / 40
print(f"pipelined: host loop {t_host*1e3:.3f} ms per batch (without the final drain), wall {t_all*1e3:.3f} ms per batch")
pr = cProfile.Profile()
pr.enable()
for _ in pipe.run(batch for _ in range(20)):
    pass
pr.disable()
torch.cuda.synchronize()
s_ = io.StringIO()
pstats.Stats(pr, stream=s_).sort_stats("tottime").print_stats(22)
print("\n".join(l[:150] for l in s_.getvalue().splitlines()[:45]))
