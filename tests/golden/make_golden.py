"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

    python tests/golden/make_golden.py

Needs /root/reference (absent on the GPU box -- the fixtures are what travels).
Inputs are not stored: every test regenerates them from ``pgmp_b200.synthetic``
with the kwargs in ``cases.py``.  Arrays above ``DIGEST_BYTES`` are stored as
sha256 digests of their raw bytes (index arrays and fp32 gathers are bit-exact
quantities, a digest pins them completely).
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import pgmp_b200  # noqa: E402
import pgmp_b200.synthetic as synthetic  # noqa: E402
import ref_shims  # noqa: E402
from cases import DIGEST_BYTES, GC_CASES, MPN_CASES, gc_config_for, mpn_config_for  # noqa: E402

import oracle  # noqa: E402  (only for the GAEC stand-in of the missing native solver)

GC_KEYS = ["x", "edge_attr", "edge_index", "joint_det", "joint_scores", "batch_index", "joint_tags"]
GC_TUPLE_POS = dict(x=0, edge_attr=1, edge_index=2, joint_det=7, joint_scores=11, batch_index=12, joint_tags=14)


def digest(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def pack(out, key, arr):
    arr = np.ascontiguousarray(arr)
    if arr.nbytes >= DIGEST_BYTES:
        out[key + "__sha256"] = digest(arr)
        out[key + "__shape"] = np.array(arr.shape, dtype=np.int64)
    else:
        out[key] = arr


def run_reference_gc(cg, name):
    inp_kw, cfg_over = GC_CASES[name]
    data = synthetic.synth_batch(**inp_kw)
    cfg = gc_config_for(pgmp_b200.config, cfg_over)
    t = {k: torch.from_numpy(v) for k, v in data.items()}
    gc = cg.NaiveGraphConstructor(
        scoremaps=t["scoremaps"], tagmaps=t["tagmaps"], features=t["features"], joints_gt=None,
        factor_list=None, masks=t["masks"] if cfg.MASK_CROWDS else None, device=torch.device("cpu"),
        config=cfg, testing=True, heatmaps=None, num_joints=inp_kw["num_joints"])
    ret = gc.construct_graph()
    assert all(ret[i] is None for i in (3, 4, 5, 6, 8, 9, 10))       # label slots at inference
    return {k: ret[GC_TUPLE_POS[k]].contiguous().numpy() for k in GC_KEYS}


def main():
    only = set(sys.argv[1:])          # optional: (re)generate just these fixtures, e.g. `mpn_pertype_hierarch_mlp`
    cg, mpn = ref_shims.load_reference()
    gc_out = {}
    for name in GC_CASES:
        if only and f"gc_{name}" not in only and not any(o.startswith("mpn_") and MPN_CASES[o[4:]][0] == name for o in only):
            continue
        res = run_reference_gc(cg, name)
        gc_out[name] = res
        if only and f"gc_{name}" not in only:
            continue
        out = {}
        for k, v in res.items():
            pack(out, k, v)
        np.savez_compressed(os.path.join(HERE, f"gc_{name}.npz"), **out)
        print(f"gc_{name}: N={len(res['joint_det'])} E={res['edge_index'].shape[1]}")

    for name, (gc_name, maker, over, seed) in MPN_CASES.items():
        if only and f"mpn_{name}" not in only:
            continue
        cfg = mpn_config_for(pgmp_b200.config, maker, over)
        model = mpn.NodeClassificationMPNSimple(cfg).eval()
        synthetic.synth_mpn_state_dict(model, seed)
        g = gc_out[gc_name]
        with torch.no_grad():
            pe, pn, pc, tag = model(torch.from_numpy(g["x"]), torch.from_numpy(g["edge_attr"]),
                                    torch.from_numpy(g["edge_index"]),
                                    node_types=torch.from_numpy(g["joint_det"][:, 2]))
        assert tag == [None]
        out = {"n_edge": np.int64(len(pe)), "n_node": np.int64(len(pn)), "n_class": np.int64(len(pc)),
               "state_sha256": digest(np.concatenate([v.numpy().astype(np.float32).ravel()
                                                      for k, v in sorted(model.state_dict().items())
                                                      if not k.endswith("num_batches_tracked")]))}
        for i, a in enumerate(pe):
            out[f"edge_{i}"] = a.numpy()
        for i, a in enumerate(pn):
            out[f"node_{i}"] = a.numpy()
        for i, a in enumerate(pc):
            out[f"class_{i}"] = a.numpy()
        np.savez_compressed(os.path.join(HERE, f"mpn_{name}.npz"), **out)
        print(f"mpn_{name}: edge |max| {np.abs(pe[-1].numpy()).max():.3f} node |max| {np.abs(pn[-1].numpy()).max():.3f}")

    if only:
        return
    # grouping tail: the reference's own weight plumbing + person assembly around OUR GAEC restatement
    pred_to_person, subgraph = ref_shims.load_reference_grouping(oracle.grouping.gaec)
    for name, gc_name, seed in [("group_knn_small", "knn_small", 0), ("group_fully_small", "fully_small", 1),
                                ("group_crowdpose", "crowdpose", 2)]:
        g = gc_out[gc_name]
        logits = synthetic.synth_group_logits(g["joint_det"], g["batch_index"], g["edge_index"],
                                              num_joints=GC_CASES[gc_name][0]["num_joints"], seed=seed)
        out = {}
        for b in np.unique(g["batch_index"]):
            sub = synthetic.image_subgraph(g, logits, int(b))
            jd = torch.from_numpy(sub["joint_det"])
            p_node = torch.from_numpy(sub["node_logits"]).sigmoid()
            p_edge = torch.from_numpy(sub["edge_logits"]).sigmoid()
            p_cls = torch.from_numpy(sub["class_logits"]).softmax(dim=1)
            ei = torch.from_numpy(sub["edge_index"])
            keep = p_node > 0.1                                           # Utils.py:1450
            ei_k, pe_k = subgraph(keep, ei, p_edge)                       # Utils.py:1451
            persons, mutants, labels = pred_to_person(jd, p_node, ei_k, pe_k, p_cls, "GAEC", sub["num_joints"])
            out[f"persons_{b}"] = np.asarray(persons, dtype=np.float64)
            out[f"labels_{b}"] = np.asarray(labels, dtype=np.int64)
            out[f"mutant_{b}"] = np.bool_(mutants)
            print(f"{name}[{b}]: persons {np.asarray(persons).shape} components {labels.max() + 1}")
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)


if __name__ == "__main__":
    main()
