"""Training-time label construction of the graph constructor (``joints_gt`` given).

Reference: ``src/graph_constructor/ConstructGraph.py`` -- the label branch of ``construct_graph`` (:104-168), the
matchings ``_construct_edge_labels_4`` (:626-686) and ``_construct_edge_labels_6`` (:769-942), ``match_cc``
(:1096-1134), ``create_loss_mask`` (:1137-1158) and the node dropout (:152-168).

Split of the work: the detection-to-ground-truth MATCHING of an image is a similarity matrix of a few hundred ground
truth joints x candidates and one or two linear sum assignments.  It runs on the host with the very operations the
reference uses (float32 torch arithmetic, ``scipy.optimize.linear_sum_assignment``) so that the assignment is the
reference's bit for bit; it needs ONE device-to-host copy of the candidates per batch.  Everything per EDGE
(``match_cc``, ``create_loss_mask``, the empty-image rule, node dropout and the re-indexing it entails) runs on the
device over the whole batch at once -- integer compares and gathers, exact by construction.
"""

import numpy as np
import torch

SUPPORTED_METHODS = (4, 6)


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


def build_labels(gc, joint_det, edge_index, batch_index, nodes_per_image):
    """Labels of the whole batch.  ``gc``: the graph constructor (config + ``joints_gt`` / ``factor_list``); the graph
    tensors are the device outputs of the emit step.  Returns a dict with the 15-tuple's label slots."""
    dev = joint_det.device
    method = gc.edge_label_method
    B = len(nodes_per_image)
    det_h = joint_det.cpu()                                  # the one host copy the matching needs
    gt_h, fac_h = gc.joints_gt.detach().cpu(), gc.factor_list.detach().cpu()
    H, W = gc.scoremaps.shape[2], gc.scoremaps.shape[3]
    N = joint_det.shape[0]
    person_h = torch.full((N,), -1, dtype=torch.int64)
    class_h = torch.zeros(N, dtype=torch.int64)
    label_h = torch.zeros(N, dtype=torch.float32)
    amb_h = torch.zeros(N, dtype=torch.bool)
    offs = np.concatenate([[0], np.cumsum(np.asarray(nodes_per_image, dtype=np.int64))])

    def one(b):
        return match_image(det_h[offs[b]:offs[b + 1]], gt_h[b], fac_h[b], method, max(H, W), gc.matching_radius,
                           gc.inclusion_radius, gc.include_neighbouring_keypoints)
    # the images are independent and torch / scipy release the GIL: a few worker threads, each running its tiny tensor
    # operations single-threaded (intra-op threads on 50 k-element tensors only get in each other's way: measured 2x)
    import os
    from concurrent.futures import ThreadPoolExecutor
    intra = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        workers = min(B, 8, max(2, (os.cpu_count() or 4) // 2))
        if B > 1:
            with ThreadPoolExecutor(max_workers=workers) as pool:
                matched = list(pool.map(one, range(B)))
        else:
            matched = [one(0)]
    finally:
        torch.set_num_threads(intra)
    off = 0
    for b in range(B):
        n = int(nodes_per_image[b])
        nodes, persons, joints, ambiguous = matched[b]
        person_h[off + nodes] = persons.long()
        class_h[off + nodes] = joints
        label_h[off + nodes] = 1.0
        if ambiguous is not None:
            amb_h[off:off + n] = torch.from_numpy(ambiguous)
        off += n
    person, node_labels, amb = person_h.to(dev), label_h.to(dev), amb_h.to(dev)
    src, dst = edge_index[0], edge_index[1]
    # match_cc: an edge is positive iff both ends are matched to joints of the same person (unmatched ends never agree)
    ps, pd = person[src], person[dst]
    edge_labels = ((ps == pd) & (ps >= 0)).float()
    # create_loss_mask: edges at ambiguous candidates leave the loss; an image without a positive edge leaves it whole
    label_mask = torch.logical_not(amb[src] | amb[dst]).float()
    has_pos = (torch.zeros(B, dtype=torch.float32, device=dev).index_add_(0, batch_index[src], edge_labels) > 0).float()
    label_mask = label_mask * has_pos[batch_index[src]]
    node_mask = torch.logical_not(amb).float()
    out = dict(edge_labels=edge_labels, node_labels=node_labels, node_persons=person, label_mask=label_mask,
               label_mask_node=node_mask, node_classes=None, class_mask=None)
    if method == 6:
        node_classes = class_h.to(dev)
        class_mask = node_labels * node_mask
        if gc.with_background_class:                        # :931-933
            node_classes = torch.where(node_labels != 1.0, torch.full_like(node_classes, gc.num_joints), node_classes)
            class_mask = torch.ones_like(class_mask)
        out.update(node_classes=node_classes, class_mask=class_mask)
    return out


def node_dropout(p_drop, x, edge_attr, edge_index, joint_det, joint_scores, batch_index, joint_tags, labels, num_images):
    """Node dropout (:152-168): every POSITIVE node is removed with probability ``p_drop`` (negatives always stay), edges at
    removed nodes go, node ids are re-numbered; over the whole batch at once.  The random draw comes from torch's device
    generator (the reference draws per image, so the streams differ; the distribution is the same)."""
    node_labels = labels["node_labels"]
    rnd = torch.bernoulli(torch.ones_like(node_labels) * p_drop)
    keep = (rnd * node_labels) == 0.0
    ekeep = keep[edge_index[0]] & keep[edge_index[1]]
    new_id = torch.cumsum(keep.long(), 0) - 1
    edge_index = new_id[edge_index[:, ekeep]]
    take = lambda t, m: t[m] if t is not None else None
    out = dict(labels)
    for k in ("node_labels", "label_mask_node", "node_classes", "node_persons", "class_mask"):
        out[k] = take(labels[k], keep)
    for k in ("edge_labels", "label_mask"):
        out[k] = take(labels[k], ekeep)
    batch_index = batch_index[keep]
    nodes_per_image = torch.bincount(batch_index, minlength=num_images)
    return (x[keep], edge_attr[ekeep], edge_index, joint_det[keep], joint_scores[keep], batch_index,
            take(joint_tags, keep), out, nodes_per_image)
