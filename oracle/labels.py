"""Test infrastructure: per-image restatement of the training-time detection-to-ground-truth matching with the
reference's own operations (float32 torch arithmetic, ``scipy.optimize.linear_sum_assignment``) --
``_construct_edge_labels_4`` (ConstructGraph.py:626-686), ``_construct_edge_labels_6`` (:769-942), ``USE_NEIGHBOURS``
(:704-727, :890-911).  Pinned to the unmodified reference through ``tests/golden/labels_*.npz``; the product
(``pgmp_b200.graph_constructor.labels`` -> ``pgmp_match_labels``, csrc/match.cu) is checked against it.  Only tests may
import this module."""
import numpy as np
import torch


def _similarity(det, gt, factors, clamp_max, floor):
    """OKS-like similarity of every annotated joint to every candidate (:773-785): ``exp(-d^2 / factor)``, float32.
    ``floor``: the smallest radius the caller thresholds the matrix with.  Far-apart pairs give subnormal results, which
    cost the CPU microcode traps (measured: 2 ms of a 5.6 ms image); when every value below ``floor`` is zeroed anyway the
    exponent is clamped at -80 first -- the values that survive the threshold keep their exact bits."""
    person_idx, joint_idx = gt[:, :, 2].nonzero(as_tuple=True)
    pos = gt[person_idx, joint_idx, :2].unsqueeze(1).round().float().clamp(0, clamp_max)
    dist = (pos - det[:, :2].float()).pow(2).sum(dim=2)
    arg = -dist / factors[person_idx, joint_idx][:, None]
    sim = torch.exp(arg.clamp_(min=-80.0) if floor > 1e-30 else arg)
    other_type = torch.logical_not(torch.eq(joint_idx.unsqueeze(1), det[:, 2]))
    return person_idx, joint_idx, sim, other_type


def _assign(cost):
    from scipy.optimize import linear_sum_assignment
    return linear_sum_assignment(cost, maximize=True)


def _neighbours(cost, rows, cols, num_gt, inclusion_radius):
    """``USE_NEIGHBOURS``: further candidates within the inclusion radius of a matched joint; candidates claimed by more
    than one joint are ambiguous and leave the loss (:704-727, :890-911).  ``cost`` is modified in place."""
    cost[cost < inclusion_radius] = 0.0
    cost[:, cols] = 0.0
    ambiguous = (cost != 0.0).sum(axis=0) > 1.0
    cost[:, ambiguous] = 0.0
    r2, _ = np.nonzero(cost)
    for r in set(r2.tolist()) - set(rows.tolist()):          # joints without a match of their own take no neighbours
        cost[r] = 0.0
    r2, c2 = np.nonzero(cost)
    lookup = np.full(num_gt, -1, dtype=np.int64)
    lookup[rows] = np.arange(len(rows), dtype=np.int64)
    return lookup[r2], c2, ambiguous


def match_image(det, gt, factors, method, clamp_max, matching_radius, inclusion_radius, use_neighbours):
    """One image: ``det [n, 3]`` int64 (x, y, type), ``gt [P, J, 3]``, ``factors [P, J]`` (CPU tensors).
    Returns ``(nodes, persons, joints, ambiguous)``: the matched candidates, the person / joint type of the ground-truth
    joint each is matched to, and the boolean ambiguity mask over the candidates (``None`` without ``USE_NEIGHBOURS``)."""
    floor = min(matching_radius, inclusion_radius) if use_neighbours else matching_radius
    person_idx, joint_idx, sim, other_type = _similarity(det, gt, factors, clamp_max, floor)
    num_gt = len(person_idx)
    if method == 4:                                          # same-type matches only (:642-652)
        sim[other_type] = 0.0
        sim[sim < matching_radius] = 0.0
        cost = sim.numpy()
        rows, cols = _assign(cost)
        keep = cost[rows, cols] != 0.0
        rows, cols = rows[keep], cols[keep]
        neigh_cost = cost
    else:                                                    # 6: same type first, any other type as a fill-in (:811-830)
        same, diff = sim.clone(), sim.clone()
        same[other_type] = 0.0
        same[same < matching_radius] = 0.0
        diff[torch.logical_not(other_type)] = 0.0
        diff[diff < matching_radius] = 0.0
        cost_same, cost_diff = same.numpy(), diff.numpy()
        sol_same, sol_diff = _assign(cost_same), _assign(cost_diff)
        rows, cols = sol_same
        fill_in = np.logical_not(cost_same[rows, cols] != 0.0)
        cols[fill_in] = sol_diff[1][fill_in]
        keep = cost_diff[sol_diff] + cost_same[sol_same] != 0.0
        rows, cols = rows[keep], cols[keep]
        neigh_cost = sim.numpy()
    persons, joints = person_idx[rows], joint_idx[rows]
    nodes = torch.from_numpy(np.ascontiguousarray(cols))
    ambiguous = None
    if use_neighbours:
        r2, c2, ambiguous = _neighbours(neigh_cost, rows, cols, num_gt, inclusion_radius)
        nodes = torch.cat([nodes, torch.from_numpy(np.ascontiguousarray(c2))])
        persons = torch.cat([persons, persons[torch.from_numpy(r2)]])
        joints = torch.cat([joints, joints[torch.from_numpy(r2)]])
    return nodes, persons, joints, ambiguous
