"""Data-parallel plumbing: images are independent units of the grouping path (no cross-image edges,
``ConstructGraph.py:58, 206-231``), so one process per GPU works on a contiguous shard of the batch and
inference needs no collective.  These helpers are the host-side logic of that scheme (``bench.py`` and
multi-GPU callers use them); ``torch.distributed`` is only used for the barrier / timing reductions and for
optionally gathering per-rank graphs back into one batch-ordered graph.
"""

import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous block split of ``total`` images: rank r owns ``[start, stop)``; sizes differ by at most 1."""
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def concat_graphs(parts):
    """Concatenate per-shard graphs (dicts with ``x, edge_attr, edge_index, joint_det, joint_scores,
    batch_index, joint_tags``) in shard order exactly like the reference batches images
    (``ConstructGraph.py:206-231``): node ids of ``edge_index`` and ``batch_index`` are offset by the
    nodes / images that precede the shard."""
    out = {k: [] for k in ("x", "edge_attr", "edge_index", "joint_det", "joint_scores", "batch_index", "joint_tags")}
    node_off, img_off = 0, 0
    for p in parts:
        n = p["joint_det"].shape[0]
        out["x"].append(p["x"])
        out["edge_attr"].append(p["edge_attr"])
        out["edge_index"].append(p["edge_index"] + node_off)
        out["joint_det"].append(p["joint_det"])
        out["joint_scores"].append(p["joint_scores"])
        out["batch_index"].append(p["batch_index"] + img_off)
        out["joint_tags"].append(p["joint_tags"])
        node_off += n
        img_off += int(p["num_images"])
    cat = {k: torch.cat(v, 1 if k == "edge_index" else 0) for k, v in out.items()}
    cat["num_images"] = img_off
    return cat


def max_over_ranks(value, device="cpu"):
    """Max of a python float over the process group (device-side timings are reported as the max over ranks)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu"):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_graphs(part):
    """All-gather every rank's graph (CPU or CUDA tensors) and return the batch-ordered concatenation."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return concat_graphs([part])
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in part.items()})
    return concat_graphs(parts)


def allreduce_gradients(params, group=None, average=True):
    """The training config's one collective: the gradients of the MPN (+ ``feature_gather``)
    parameters -- 0.4 MB agnostic / 1.5 MB per-type, fp32 -- are packed into ONE bucket, all-reduced
    (NCCL over NVLink / NVSwitch on the GPU box, gloo in the CPU tests) and unpacked in place; ``average``
    divides by the world size (what ``DistributedDataParallel`` would do for the reference's ``train.py``).
    Parameters without a gradient contribute zeros so that every rank reduces the same bucket layout.
    Returns the number of bytes reduced (0 without a process group)."""
    params = [p for p in params if p.requires_grad]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1 or not params:
        return 0
    world = dist.get_world_size(group)
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= world
    off = 0
    for p in params:
        n = p.numel()
        if p.grad is None:
            p.grad = flat[off:off + n].view_as(p).clone()
        else:
            p.grad.copy_(flat[off:off + n].view_as(p))
        off += n
    return flat.numel() * flat.element_size()
