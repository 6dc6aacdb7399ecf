"""Grouping tail of the hot path: logits -> persons.

Batched CUDA equivalent of ``src/valid.py:109-122`` + ``pred_to_ann`` up to ``pred_to_person``
(``src/Utils/Utils.py:1445-1457, 499-514``), ``cluster_graph`` with ``CC_METHOD == "GAEC"``
(``src/Utils/correlation_clustering/correlation_clustering_utils.py``) and
``graph_cluster_to_persons`` (``Utils.py:672-743``).  The reference does this per image on the host
with dense N x N numpy matrices and a native solver that is missing from its tree; here one CTA per
image runs the whole tail on the device (``csrc/group.cu``).
"""

import numpy as np
import torch

from .. import _native as nv
from .transformations import (gen_ann_format, gen_ann_format_correct, gen_ann_format_mean, persons_to_ann,  # noqa: F401
                              reverse_affine_map, reverse_affine_map_points)


class PendingGroups:
    """Handle of a grouping launch (``group_persons_async``): ``result()`` waits for the device and splits the packed
    outputs per image."""

    def __init__(self, B, max_persons, labels, node_off_h, small_h, persons_h, event):
        self.B, self.max_persons, self.labels, self.node_off_h = B, max_persons, labels, node_off_h
        self.small_h, self.persons_h, self.event = small_h, persons_h, event

    def result(self):
        self.event.synchronize()
        small = self.small_h.numpy()                 # [5, B]: components, kept edges, persons, mutants, detector ok
        persons = self.persons_h.numpy()
        out = []
        for b in range(self.B):
            nk, npers, mut, ok = int(small[1, b]), int(small[2, b]), int(small[3, b]), int(small[4, b])
            if npers > self.max_persons:
                raise RuntimeError("more than max_persons=%d persons in image %d" % (self.max_persons, b))
            if nk <= 0 or ok < 1:
                out.append(None)
                continue
            out.append((persons[b, :npers].copy() if npers else np.array([]), bool(mut),
                        self.labels[self.node_off_h[b]:self.node_off_h[b + 1]]))
        return out


def group_persons_async(joint_det, node_logits, edge_index, edge_logits, class_logits, batch_index, num_joints,
                        node_threshold=0.1, cc_method="GAEC", max_persons=None, detector_scores=None,
                        nodes_per_image=None, edges_per_image=None, stream=None):
    """``group_persons`` without the wait: launches on ``stream`` (default: the current one), copies the packed results to
    pinned host memory behind an event and returns a ``PendingGroups``.

    ``nodes_per_image`` / ``edges_per_image``: the per-image counts the graph constructor already holds on the host
    (``gc.num_nodes_per_image`` / ``gc.num_edges_per_image``).  They fix the batch size -- trailing images without
    candidates stay in the result as ``None`` -- and spare the device-side counting and its host reads; without them the
    batch size is ``batch_index[-1] + 1`` and the counts are read back once.
    """
    if cc_method not in nv.CC_METHODS:
        raise NotImplementedError("CC_METHOD=%r (GAEC, the reference default, threshold and greedy are in scope; KL / MUT need "
                                  "the reference's missing native solver)" % (cc_method,))
    nv.require_cuda(joint_det, "joint_det", torch.int64)
    nv.require_cuda(node_logits, "node_logits", torch.float32)
    nv.require_cuda(edge_index, "edge_index", torch.int64)
    nv.require_cuda(edge_logits, "edge_logits", torch.float32)
    nv.require_cuda(batch_index, "batch_index", torch.int64)
    dev = joint_det.device
    N, E = joint_det.shape[0], edge_index.shape[1]
    lib = nv.lib()
    if nodes_per_image is not None:
        nodes_h = torch.as_tensor(nodes_per_image, dtype=torch.int64).cpu()
        B = int(nodes_h.numel())
        if edges_per_image is not None:
            edges_h = torch.as_tensor(edges_per_image, dtype=torch.int64).cpu()
        else:
            edges_h = (torch.bincount(batch_index[edge_index[0]], minlength=B) if E else torch.zeros(B, dtype=torch.int64)).cpu()
    else:
        if N == 0:
            return PendingGroups(0, 0, None, [0], torch.zeros((5, 0), dtype=torch.int32), torch.zeros(0), torch.cuda.Event())
        B = int(batch_index[-1].item()) + 1
        nodes_h = torch.bincount(batch_index, minlength=B).cpu()
        edges_h = (torch.bincount(batch_index[edge_index[0]], minlength=B) if E else torch.zeros(B, dtype=torch.int64)).cpu()
    if int(nodes_h.sum()) != N or int(edges_h.sum()) != E:
        raise ValueError("nodes_per_image / edges_per_image do not add up to the graph's size")
    off_h = torch.zeros((2, B + 1), dtype=torch.int64, pin_memory=True)
    off_h[0, 1:] = nodes_h.cumsum(0)
    off_h[1, 1:] = edges_h.cumsum(0)
    max_nodes = max(int(nodes_h.max()) if B else 0, 1)
    J = int(num_joints)
    max_persons = int(max_persons or max(1, max_nodes // 2))
    with torch.cuda.device(dev), torch.cuda.stream(stream if stream is not None else torch.cuda.current_stream(dev)):
        off = off_h.to(dev, non_blocking=True)
        labels = torch.empty(N, dtype=torch.int64, device=dev)
        small = torch.zeros((5, B), dtype=torch.int32, device=dev)
        persons = torch.zeros((B, max_persons, J, 3), dtype=torch.float64, device=dev)
        jd = joint_det.contiguous()
        nl = node_logits.detach().reshape(-1).contiguous()
        el = edge_logits.detach().reshape(-1).contiguous()
        ei = edge_index.contiguous()
        cl = class_logits.detach().contiguous() if class_logits is not None else None
        if detector_scores is not None:                            # Utils.py:1448-1449
            small[4] = torch.zeros(B, dtype=torch.int64, device=dev).index_add_(
                0, batch_index, (detector_scores > 0.1).long()).clamp_(max=1).int()
        else:
            small[4] = 1
        if N > 0:
            p = nv.GroupParams(batch=B, num_joints=J, num_nodes=N, num_edges=E, node_threshold=float(node_threshold),
                               cc_method=nv.CC_METHODS[cc_method], edge_threshold=0.8,
                               node_offsets=off[0].data_ptr(), edge_offsets=off[1].data_ptr(), edge_index=ei.data_ptr(),
                               joint_det=jd.data_ptr(), node_logits=nl.data_ptr(), edge_logits=el.data_ptr(),
                               class_logits=cl.data_ptr() if cl is not None else None, person_labels=labels.data_ptr(),
                               num_components=small[0].data_ptr(), num_kept_edges=small[1].data_ptr(), max_persons=max_persons,
                               max_nodes_per_image=max_nodes, persons=persons.data_ptr(), num_persons=small[2].data_ptr(),
                               mutants=small[3].data_ptr())
            ws_bytes = int(lib.pgmp_group_workspace_bytes(p))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            p.workspace, p.workspace_bytes = ws.data_ptr(), ws_bytes
            nv.check(lib.pgmp_group_persons(p, nv.current_stream()))
        small_h = torch.empty(small.shape, dtype=torch.int32, pin_memory=True)
        persons_h = torch.empty(persons.shape, dtype=torch.float64, pin_memory=True)
        small_h.copy_(small, non_blocking=True)
        persons_h.copy_(persons, non_blocking=True)
        event = torch.cuda.Event()
        event.record()
        if stream is not None:                                     # inputs produced on the caller's stream, read on this one
            for t_ in (jd, nl, el, ei, batch_index) + ((cl,) if cl is not None else ()):
                t_.record_stream(stream)
    return PendingGroups(B, max_persons, labels, off_h[0].tolist(), small_h, persons_h, event)


def group_persons(joint_det, node_logits, edge_index, edge_logits, class_logits, batch_index, num_joints,
                  node_threshold=0.1, cc_method="GAEC", max_persons=None, detector_scores=None,
                  nodes_per_image=None, edges_per_image=None):
    """Returns one entry per image of the batch: ``None`` where the reference's ``pred_to_ann`` returns
    ``None`` before grouping (no detector score > 0.1, Utils.py:1448-1449; no edge between kept nodes,
    :1452,1457) else ``(persons [P, J, 3] float64 ndarray, mutants bool, person_labels [N_b] int64 tensor)``
    as ``pred_to_person`` does.

    ``node_logits`` / ``edge_logits`` / ``class_logits`` are the MPN's last predictions (logits: the sigmoid /
    softmax of valid.py:109-111 are applied on the device); ``edge_index`` holds global node ids with the edges of
    an image contiguous, as ``construct_graph`` returns them.  Pass ``nodes_per_image`` / ``edges_per_image`` (the graph
    constructor's host-side counts) to keep trailing empty images in the result; one packed device-to-host copy.
    """
    return group_persons_async(joint_det, node_logits, edge_index, edge_logits, class_logits, batch_index, num_joints,
                               node_threshold=node_threshold, cc_method=cc_method, max_persons=max_persons,
                               detector_scores=detector_scores, nodes_per_image=nodes_per_image,
                               edges_per_image=edges_per_image).result()


def refine_persons(scoremaps, tags, persons, with_refine=True, adjustment=True):
    """The pose-assembly tail of ``pred_to_ann`` (``src/Utils/Utils.py:1472-1477``) for a whole batch: ``refine``
    (:1026-1104) places the joints a person is missing where the heatmap is high and the tag is close to the person's
    mean tag, ``adjust`` (:917-936) adds the quarter-pixel offsets.  ``scoremaps [B, J, H, W]`` and ``tags [B, J, H, W]``
    (or ``[B, J, H, W, T]``) are CUDA float32 tensors, ``persons`` is a list with one ``[P, J, 3]`` float64 array
    (x, y, score; e.g. ``group_persons(...)[b][0]`` after ``fill_mean``) or ``None`` per image.  Returns the updated list.
    """
    nv.require_cuda(scoremaps, "scoremaps", torch.float32)
    nv.require_cuda(tags, "tags", torch.float32)
    B, J, H, W = scoremaps.shape
    if len(persons) != B or tags.shape[:4] != scoremaps.shape:
        raise ValueError("one persons entry per image; tags must be [B, J, H, W(, T)]")
    T = tags.shape[4] if tags.dim() == 5 else 1
    counts = [0 if q is None or len(q) == 0 else int(q.shape[0]) for q in persons]
    pmax = max(counts + [1])
    host = np.zeros((B, pmax, J, 3), dtype=np.float64)
    for b, q in enumerate(persons):
        if counts[b]:
            host[b, :counts[b]] = q
    dev = scoremaps.device
    kp = torch.from_numpy(host).to(dev)
    npers = torch.tensor(counts, dtype=torch.int32, device=dev)
    sm, tg = scoremaps.contiguous(), tags.contiguous()
    p = nv.RefineParams(batch=B, num_joints=J, height=H, width=W, tag_dim=T, max_persons=pmax, do_refine=int(bool(with_refine)),
                        do_adjust=int(bool(adjustment)), scoremaps=sm.data_ptr(), tags=tg.data_ptr(), persons=kp.data_ptr(),
                        num_persons=npers.data_ptr())
    lib = nv.lib()
    with torch.cuda.device(dev):
        ws_bytes = int(lib.pgmp_refine_workspace_bytes(p))
        ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)
        p.workspace, p.workspace_bytes = ws.data_ptr(), ws.numel()
        nv.check(lib.pgmp_refine_persons(p, nv.current_stream()))
    out_h = kp.cpu().numpy()
    return [None if q is None else (out_h[b, :counts[b]].copy() if counts[b] else q) for b, q in enumerate(persons)]


def filter_and_fill(persons, with_filter=False, fill_mean=True):
    """The host steps of ``pred_to_ann`` between the grouping and ``refine`` (``src/Utils/Utils.py:1463-1471``) for one
    image: ``with_filter`` keeps the persons whose best joint score exceeds 0.25 (returns ``None`` when none is left, as
    the reference returns no annotation), ``fill_mean`` moves every missing joint (score 0) to the mean position of the
    person's detected joints.  ``persons`` is a ``[P, J, 3]`` float64 array (x, y, score); a copy is returned."""
    p = np.array(persons, dtype=np.float64, copy=True)
    if p.ndim != 3:                                    # Utils.py:1458-1460: no person
        return None
    if with_filter:
        p = p[p[:, :, 2].max(axis=1) > 0.25]
        if p.shape[0] == 0:
            return None
    if fill_mean:
        for i in range(len(p)):
            missing = p[i, :, 2] == 0
            p[i, missing, :2] = p[i, ~missing, :2].mean(axis=0)
    return p


def persons_from_groups(scoremaps, tags, groups, with_refine=True, adjustment=True, with_filter=False, fill_mean=True):
    """``pred_to_ann`` from the grouping to the final keypoints (``Utils.py:1458-1477``) for a whole batch:
    ``groups`` is what ``group_persons`` returns; per image ``filter_and_fill`` on the host (a few dozen floats), then
    ``refine`` / ``adjust`` for all images in one device call.  Returns one ``[P, J, 3]`` float64 array or ``None`` per image
    (the input of ``reverse_affine_map``, :1478)."""
    persons = []
    for g in groups:
        persons.append(None if g is None else filter_and_fill(g[0], with_filter, fill_mean))
    if not (with_refine or adjustment) or all(q is None for q in persons):
        return persons
    return refine_persons(scoremaps, tags, persons, with_refine=with_refine, adjustment=adjustment)
