import os, sys
sys.path[:0] = ['.', 'tests', 'tests/golden']
import numpy as np, torch
import importlib.util
spec = importlib.util.spec_from_file_location("tgt", "tests/test_gpu_train.py")
m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
import oracle.mpn_train as T
import pgmp_b200
from cases import mpn_config_for
for name in sys.argv[1:]:
    gc_name, over, seed, _ = m.VARIANTS[name]
    g = m.graph_for(gc_name)
    cfg = mpn_config_for(pgmp_b200.config, "agnostic_mpn_config", over)
    model, sd0, x, pe, pn, pc, coeffs, loss = m.run_cuda(cfg, seed, g)
    out = T.loss_and_gradients(sd0, cfg, g["x"], g["edge_attr"], g["edge_index"], g["joint_det"][:, 2], coeffs)
    ograds = out[5]
    params = dict(model.named_parameters())
    print("==", name)
    for pn_ in ("classification.0.weight", "classification.0.bias", "node_classification.0.weight", "mpn_node_cls.mlp_node.0.weight", "mpn_node_cls.mlp_edge.2.weight"):
        if pn_ not in params: continue
        got = params[pn_].grad.cpu().numpy().astype(np.float64); want = ograds[pn_]
        err = np.abs(got - want)
        print(pn_, "max|want|", np.abs(want).max(), "max err", err.max())
        if err.ndim == 2:
            print("  err by out row (max):", np.round(err.max(1) / np.abs(want).max(), 4)[:64])
            print("  err by in col (max):", np.round(err.max(0) / np.abs(want).max(), 4)[:64])
        else:
            print("  err:", np.round(err / np.abs(want).max(), 4))
