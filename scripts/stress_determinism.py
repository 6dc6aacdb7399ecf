import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import pgmp_b200, pgmp_b200.synthetic as synthetic
from pgmp_b200.graph_constructor import get_graph_constructor
from pgmp_b200.Models.MessagePassingNetwork import get_mpn_model
dev="cuda:0"
for (B,J,S,K,graph) in [(32,17,512,30,"knn"),(3,17,512,30,"fully"),(1,17,512,30,"knn"),(5,14,512,60,"knn")]:
    sm = torch.from_numpy(np.stack([synthetic.synth_scoremap(b, J, S, K, persons=8 if J==17 else 20) for b in range(B)])).to(dev)
    g = torch.Generator(device=dev).manual_seed(0)
    feat = torch.randn(B,128,S,S,device=dev,generator=g); tags = torch.randn(B,J,S,S,device=dev,generator=g)
    gcfg = pgmp_b200.config.bench_gc_config(k=K, graph_type=graph)
    over = {} if J==17 else dict(NUM_JOINTS=J, EDGE_INPUT_DIM=J+2)
    mcfg = pgmp_b200.config.flagship_mpn_config(J, B200_PRECISION="tc", AUX_LOSS_STEPS=2, **over)
    if J!=17: mcfg.CLASS.OUTPUT_SIZES=[64,32,J]
    model = synthetic.synth_mpn_state_dict(get_mpn_model(mcfg), 1).eval().to(dev)
    ref=None; bad=0
    for it in range(25):
        ret = get_graph_constructor(gcfg, scoremaps=sm, tagmaps=tags, features=feat, joints_gt=None, factor_list=None, masks=None, device=dev, testing=True, heatmaps=None, num_joints=J).construct_graph()
        with torch.no_grad():
            pe,pn,pc,_ = model(ret[0],ret[1],ret[2],node_types=ret[7][:,2])
        cur=[t.clone() for t in list(pe)+list(pn)+list(pc)]+[ret[0].clone(),ret[2].clone()]
        if ref is None: ref=cur
        else:
            for a,b in zip(ref,cur):
                if not torch.equal(a,b): bad+=1
    print(B,J,graph,"mismatching tensors over 24 repeats:",bad, "finite:", all(bool(torch.isfinite(t).all()) for t in ref[:-1]))
