"""Training-time label construction of the graph constructor (``joints_gt`` given).

Reference: ``src/graph_constructor/ConstructGraph.py`` -- the label branch of ``construct_graph`` (:104-168), the
matchings ``_construct_edge_labels_4`` (:626-686) and ``_construct_edge_labels_6`` (:769-942), ``match_cc``
(:1096-1134), ``create_loss_mask`` (:1137-1158) and the node dropout (:152-168).

Split of the work: the detection-to-ground-truth MATCHING of an image is a similarity matrix of a few hundred ground
truth joints x candidates and one or two linear sum assignments.  It runs on the host, natively and for all images of
the batch in parallel (``csrc/match.cu``: ``pgmp_label_similarity_args`` computes the exponents with the reference's
float32 arithmetic, ``torch.exp`` -- the reference's own routine -- turns them into similarities, ``pgmp_match_labels``
does the masks, thresholds, assignments -- SciPy's ``linear_sum_assignment`` algorithm restated, checked against SciPy --
fill-in and neighbour rules), so that the assignment is the reference's bit for bit; it needs ONE device-to-host copy
of the candidates per batch.  The per-image torch / scipy form of the same matching is kept as test infrastructure in
``oracle/labels.py``.  Everything per EDGE
(``match_cc``, ``create_loss_mask``, the empty-image rule, node dropout and the re-indexing it entails) runs on the
device over the whole batch at once -- integer compares and gathers, exact by construction.
"""

import numpy as np
import torch

SUPPORTED_METHODS = (4, 6)


def build_labels(gc, joint_det, edge_index, batch_index, nodes_per_image):
    """See ``_build_labels``.  The host tensors here are small (the largest, the similarity matrices, take 0.3 ms on one
    thread), so torch's intra-op pool is switched off for the duration: OpenMP workers left spinning after a parallel
    region share the cores with the matcher's threads and double the time of this function (measured: 6.5 -> 3.2 ms)."""
    intra = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        return _build_labels(gc, joint_det, edge_index, batch_index, nodes_per_image)
    finally:
        torch.set_num_threads(intra)


def _build_labels(gc, joint_det, edge_index, batch_index, nodes_per_image):
    """Labels of the whole batch.  ``gc``: the graph constructor (config + ``joints_gt`` / ``factor_list``); the graph
    tensors are the device outputs of the emit step.  Returns a dict with the 15-tuple's label slots."""
    dev = joint_det.device
    method = gc.edge_label_method
    B = len(nodes_per_image)
    import ctypes as C
    import os
    from .. import _native as nv
    lib = nv.lib()
    det_h = joint_det.cpu().contiguous()                     # the one host copy the matching needs
    gt_h, fac_h = gc.joints_gt.detach().cpu(), gc.factor_list.detach().cpu()
    if fac_h.dtype != torch.float32 or not gt_h.is_floating_point():
        raise NotImplementedError("joints_gt / factor_list: float tensors with float32 factors (the reference's data pipeline)")
    vis = gt_h[:, :, :, 2] != 0
    gt32 = torch.cat([gt_h[..., :2].round().float(), vis.float().unsqueeze(-1)], dim=-1).contiguous()   # rounded in the source dtype (:774)
    fac_h = fac_h.contiguous()
    H, W = gc.scoremaps.shape[2], gc.scoremaps.shape[3]
    N = joint_det.shape[0]
    P, J = gt_h.shape[1], gt_h.shape[2]
    use_nb = bool(gc.include_neighbouring_keypoints)
    floor = min(gc.matching_radius, gc.inclusion_radius) if use_nb else gc.matching_radius
    counts = np.asarray(nodes_per_image, dtype=np.int64)
    offs = torch.from_numpy(np.concatenate([[0], np.cumsum(counts)]))
    n_max = max(int(counts.max()) if B else 0, 1)
    g_max = max(int(vis.sum(dim=(1, 2)).max()) if B else 0, 1)
    threads = min(max(B, 1), 16, os.cpu_count() or 4)
    # 1. the exponent -d^2 / factor of the similarity (:773-785) for the whole batch, natively -- float32, operation for
    #    operation the reference's arithmetic.  Far-apart pairs would give subnormal similarities, which cost the CPU
    #    microcode traps; when every value below `floor` is zeroed anyway the exponent is clamped at -80 first: the values
    #    that survive the thresholds keep their exact bits.
    arg = torch.zeros((B, g_max, n_max), dtype=torch.float32)
    num_gt = torch.zeros(B, dtype=torch.int32)
    gt_type = torch.zeros((B, g_max), dtype=torch.int32)
    gt_person = torch.zeros((B, g_max), dtype=torch.int32)
    det_type = torch.zeros((B, n_max), dtype=torch.int32)
    ap = nv.LabelArgsParams(batch=B, max_persons=P, num_joints=J, num_threads=threads, clamp_max=float(max(H, W)),
                            min_arg=-80.0 if floor > 1e-30 else float("-inf"), det=det_h.data_ptr(),
                            node_offsets=offs.data_ptr(), gt=gt32.data_ptr(), factors=fac_h.data_ptr(), max_gt=g_max,
                            max_det=n_max, arg=arg.data_ptr(), num_gt=num_gt.data_ptr(), gt_type=gt_type.data_ptr(),
                            gt_person=gt_person.data_ptr(), det_type=det_type.data_ptr())
    nv.check(lib.pgmp_label_similarity_args(C.byref(ap)))
    # 2. exp() with the reference's own routine, so that every similarity carries the reference's bits
    sim = torch.exp_(arg)
    # 3. masks, radius thresholds, the linear sum assignments, fill-in and neighbour rules, per-node labels: natively
    cap = n_max + g_max
    match_row = torch.empty((B, cap), dtype=torch.int32)
    match_col = torch.empty((B, cap), dtype=torch.int32)
    num_match = torch.zeros(B, dtype=torch.int32)
    amb_pad = torch.zeros((B, n_max), dtype=torch.uint8)
    person_h = torch.full((N,), -1, dtype=torch.int64)
    class_h = torch.zeros(N, dtype=torch.int64)
    label_h = torch.zeros(N, dtype=torch.float32)
    amb_u8 = torch.zeros(N, dtype=torch.uint8)
    num_det = torch.from_numpy(counts.astype(np.int32))
    mp = nv.MatchParams(batch=B, method=method, use_neighbours=int(use_nb), num_threads=threads,
                        matching_radius=gc.matching_radius, inclusion_radius=gc.inclusion_radius, sim=sim.data_ptr(),
                        sim_stride_b=sim.stride(0), sim_stride_g=sim.stride(1), max_gt=g_max, max_det=n_max,
                        num_gt=num_gt.data_ptr(), num_det=num_det.data_ptr(), gt_type=gt_type.data_ptr(),
                        det_type=det_type.data_ptr(), cap=cap, match_row=match_row.data_ptr(),
                        match_col=match_col.data_ptr(), num_match=num_match.data_ptr(), ambiguous=amb_pad.data_ptr(),
                        node_offsets=offs.data_ptr(), gt_person=gt_person.data_ptr(), node_person=person_h.data_ptr(),
                        node_class=class_h.data_ptr(), node_label=label_h.data_ptr(), node_ambiguous=amb_u8.data_ptr())
    nv.check(lib.pgmp_match_labels(C.byref(mp)))
    amb_h = amb_u8.bool()
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
