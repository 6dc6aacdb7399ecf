"""Golden vectors of the grouping tail with CC_METHOD = "threshold" (Utils.py:508-509) and "greedy" (:517-626), produced
by the reference's own ``pred_to_person`` / ``graph_cluster_to_persons`` / ``greedy_person_construction`` -- no stand-in
is involved on these paths.

    python tests/golden/make_golden_threshold.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import pgmp_b200.synthetic as synthetic  # noqa: E402
import ref_shims  # noqa: E402
from cases import GC_CASES  # noqa: E402
from make_golden import run_reference_gc  # noqa: E402


def main():
    cg, _ = ref_shims.load_reference()
    pred_to_person, subgraph = ref_shims.load_reference_grouping(None)
    for name, gc_name, seed, method in [("group_threshold_knn_small", "knn_small", 0, "threshold"),
                                        ("group_threshold_crowdpose", "crowdpose", 2, "threshold"),
                                        ("group_greedy_knn_small", "knn_small", 0, "greedy"),
                                        ("group_greedy_fully_small", "fully_small", 1, "greedy"),
                                        ("group_greedy_crowdpose", "crowdpose", 2, "greedy")]:
        g = run_reference_gc(cg, gc_name)
        logits = synthetic.synth_group_logits(g["joint_det"], g["batch_index"], g["edge_index"],
                                              num_joints=GC_CASES[gc_name][0]["num_joints"], seed=seed)
        out = {}
        for b in np.unique(g["batch_index"]):
            sub = synthetic.image_subgraph(g, logits, int(b))
            jd = torch.from_numpy(sub["joint_det"])
            p_node = torch.from_numpy(sub["node_logits"]).sigmoid()
            p_edge = torch.from_numpy(sub["edge_logits"]).sigmoid()
            p_cls = torch.from_numpy(sub["class_logits"]).softmax(dim=1)
            keep = p_node > 0.1                                           # Utils.py:1450
            ei_k, pe_k = subgraph(keep, torch.from_numpy(sub["edge_index"]), p_edge)   # Utils.py:1451
            persons, mutants, labels = pred_to_person(jd.clone(), p_node, ei_k, pe_k, p_cls, method, sub["num_joints"])
            out[f"persons_{b}"] = np.asarray(persons, dtype=np.float64)
            out[f"labels_{b}"] = np.asarray(labels, dtype=np.int64)
            out[f"mutant_{b}"] = np.bool_(mutants)
            print(f"{name}[{b}]: persons {np.asarray(persons).shape} clusters {len(np.unique(labels))}")
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)


if __name__ == "__main__":
    main()
