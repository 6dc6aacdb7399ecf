"""Message-passing-network oracle (numpy, fp32).  TEST INFRASTRUCTURE -- see oracle/__init__.py.

Follows ``src/Models/MessagePassingNetwork/{layers,NodeClassificationMPNSimple,utils}.py``.
Weights arrive as a ``state_dict``-like mapping ``name -> ndarray`` with the
reference's parameter names, eval-mode BatchNorm.
"""

import numpy as np

BN_EPS = np.float32(1e-5)          # nn.BatchNorm1d default (layers.py:14)
SOFTMAX_EPS = np.float32(1e-12)    # torch_scatter.composite.scatter_softmax eps


def relu(v):
    return np.maximum(v, np.float32(0))


def linear(sd, name, v):
    return (v @ sd[name + ".weight"].T + sd[name + ".bias"]).astype(np.float32)


# ---------------------------------------------------------------- _make_mlp
def mlp_spec(hidden_dims, bn=False, end_with_relu=False):
    """Module layout of ``_make_mlp`` (layers.py:8-29) as a list of
    ("linear"|"relu"|"bn", sequential_index)."""
    ops, idx = [], 0

    def push(kind):
        nonlocal idx
        ops.append((kind, idx))
        idx += 1

    push("linear")                                   # layers.py:10
    if len(hidden_dims) != 1:
        push("relu")                                 # layers.py:11-12
    if bn and len(hidden_dims) != 1:
        push("bn")                                   # layers.py:13-14
    for i in range(1, len(hidden_dims)):
        push("linear")                               # layers.py:16
        if i != len(hidden_dims) - 1:
            push("relu")                             # layers.py:20-21
            if bn:
                push("bn")                           # layers.py:22-23
    if end_with_relu:
        push("relu")                                 # layers.py:24-25
        if bn:
            push("bn")                               # layers.py:26-27
    return ops


def mlp_forward(sd, prefix, hidden_dims, v, bn=False, end_with_relu=False):
    for kind, i in mlp_spec(hidden_dims, bn, end_with_relu):
        name = f"{prefix}.{i}"
        if kind == "linear":
            v = linear(sd, name, v)
        elif kind == "relu":
            v = relu(v)
        else:  # eval-mode BatchNorm1d
            inv = np.float32(1) / np.sqrt(sd[name + ".running_var"] + BN_EPS)
            v = ((v - sd[name + ".running_mean"]) * inv * sd[name + ".weight"] + sd[name + ".bias"]).astype(np.float32)
    return v


# ------------------------------------------------------------- scatter helpers
def _segments(index):
    order = np.argsort(index, kind="stable")
    sidx = index[order]
    starts = np.flatnonzero(np.r_[True, sidx[1:] != sidx[:-1]]) if len(sidx) else np.zeros(0, np.int64)
    return order, sidx, starts


def scatter(values, index, dim_size, reduce):
    """``torch_scatter.scatter(values, index, dim=0, dim_size=, reduce=)``; empty
    segments are 0 (also for max, as torch_scatter does)."""
    out = np.zeros((dim_size,) + values.shape[1:], dtype=np.float32)
    if len(index) == 0:
        return out
    order, sidx, starts = _segments(index)
    v = values[order]
    if reduce in ("add", "sum", "mean"):
        red = np.add.reduceat(v, starts, axis=0)
        if reduce == "mean":
            cnt = np.diff(np.r_[starts, len(sidx)]).astype(np.float32)
            red = red / cnt.reshape((-1,) + (1,) * (v.ndim - 1))
    elif reduce == "max":
        red = np.maximum.reduceat(v, starts, axis=0)
    else:
        raise NotImplementedError(reduce)
    out[sidx[starts]] = red
    return out


def scatter_softmax(src, index):
    """``torch_scatter.composite.scatter_softmax`` on a 1-D ``src``."""
    if len(index) == 0:
        return src.copy()
    order, sidx, starts = _segments(index)
    seg_id = np.cumsum(np.r_[0, (sidx[1:] != sidx[:-1]).astype(np.int64)])
    v = src[order]
    mx = np.maximum.reduceat(v, starts)
    ex = np.exp(v - mx[seg_id]).astype(np.float32)
    sm = np.add.reduceat(ex, starts) + SOFTMAX_EPS
    out = np.empty_like(src)
    out[order] = ex / sm[seg_id]
    return out.astype(np.float32)


# ---------------------------------------------------------------- node-type merge
def sum_node_types(node_summary, node_types):
    """MessagePassingNetwork/utils.py:6-19."""
    if node_summary == "not":
        return node_types
    if node_summary == "left_right":
        return np.array([0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8])[node_types]
    if node_summary == "per_body_part":
        return np.array([0, 0, 0, 0, 0, 1, 1, 2, 3, 2, 3, 4, 5, 4, 5, 4, 5])[node_types]
    raise NotImplementedError(node_summary)


# ---------------------------------------------------------------- the two layers
def _edge_update(sd, p, x, e, edge_index):
    src, dst = edge_index                                        # layers.py:210 (j, i)
    cat = np.concatenate([x[dst], x[src], e], 1)                 # layers.py:214: [x_i ; x_j ; e_ij]
    h = relu(linear(sd, p + ".mlp_edge.0", cat))
    return relu(linear(sd, p + ".mlp_edge.2", h))                # layers.py:171-175


def mp_layer(sd, p, x, e, edge_index, aggr, use_node_update_mlp):
    """``MPLayer.forward`` (agnostic edge MLP), layers.py:63-86."""
    src, dst = edge_index
    e_new = _edge_update(sd, p, x, e, edge_index)
    m = relu(linear(sd, p + ".mlp_node.0", np.concatenate([x[dst], e_new], 1)))   # layers.py:78-81
    out = scatter(m, dst, x.shape[0], aggr)                      # PyG default aggregate over edge_index[1]
    if use_node_update_mlp:
        out = relu(linear(sd, p + ".update_mlp.0", out))         # layers.py:83-86
    return out, e_new


def type_aware_layer(sd, p, x, e, edge_index, node_types, aggr, aggr_sub, num_types, update_type="mlp"):
    """``TypeAwareMPNLayer.forward`` (agnostic edge MLP, ``update_type == "mlp"``), layers.py:207-258."""
    src, dst = edge_index
    n = x.shape[0]
    e_new = _edge_update(sd, p, x, e, edge_index)
    src_type = node_types[src]                                   # layers.py:220
    inp = np.concatenate([x[dst], e_new], 1)                     # layers.py:223,273: [x_i ; e']
    m = np.zeros((len(src), sd[p + ".mlp_node.mlp.0.0.bias"].shape[0]), dtype=np.float32)
    for t in range(17):                                          # layers.py:271 (hard-coded 17)
        sel = src_type == t
        if sel.any():
            m[sel] = relu(linear(sd, f"{p}.mlp_node.mlp.{t}.0", inp[sel]))
    upd = np.zeros((n, num_types, m.shape[1]), dtype=np.float32)
    if aggr_sub == "None":                                       # layers.py:234-240
        for t in range(num_types):
            sel = src_type == t
            upd[:, t] = scatter(m[sel], dst[sel], n, aggr)
    elif aggr_sub in ("node_edge_attn", "node_edge_attn_per_type"):   # layers.py:242-251
        attn = linear(sd, p + ".attn_net.0", e_new)
        for t in range(num_types):
            col = 0 if aggr_sub == "node_edge_attn" else t
            sel = src_type == t
            a = scatter_softmax(attn[sel, col], dst[sel])
            upd[:, t] = scatter(m[sel] * a[:, None], dst[sel], n, "add")
    else:
        raise NotImplementedError(aggr_sub)
    if update_type == "hierarch_mlp":
        return hierarch_update_mlp(sd, p + ".update_mlp", upd, num_types), e_new
    out = relu(linear(sd, p + ".update_mlp.0", upd.reshape(n, -1)))   # layers.py:253-258
    return out, e_new


def hierarch_update_mlp(sd, p, upd, num_joints):
    """``HierarchUpdateMlp.forward`` (layers.py:109-128): body-part tree over the per-type aggregates [N, T, D]."""
    n = upd.shape[0]
    if num_joints == 17:                                             # layers.py:114-116
        order_1 = [(0, 1, 2, 3, 4), (5, 6), (7, 9), (8, 10), (11, 12), (13, 15), (14, 16)]
    else:                                                            # layers.py:117-119
        order_1 = [(0, 1), (2, 3), (4, 6), (5, 7), (8, 9), (10, 12), (11, 13)]
    order_2 = [(0, 1), (1, 2), (1, 3), (1, 4), (4, 5), (4, 6)]
    out_1 = np.stack([relu(linear(sd, f"{p}.first_layer.{i}", upd[:, list(t)].reshape(n, -1))) for i, t in enumerate(order_1)], 1)
    out_2 = np.stack([relu(linear(sd, f"{p}.second_layer.{i}", out_1[:, list(t)].reshape(n, -1))) for i, t in enumerate(order_2)], 1)
    return relu(linear(sd, p + ".final", out_2.reshape(n, -1)))


# ---------------------------------------------------------------- the model
def node_classification_mpn_forward(sd, cfg, x, edge_attr, edge_index, node_types):
    """``NodeClassificationMPNSimple.forward``, NodeClassificationMPNSimple.py:62-97.

    cfg carries the reference's ``MODEL.MPN`` attribute names.  Returns
    (preds_edge, preds_node, preds_class) as lists of arrays; ``squeeze()``
    semantics of :82,:84,:93 are reproduced.
    """
    x = np.asarray(x, dtype=np.float32)
    edge_attr = np.asarray(edge_attr, dtype=np.float32)
    node_types = sum_node_types(cfg.NODE_TYPE_SUMMARY, np.asarray(node_types))    # :64
    h = mlp_forward(sd, "node_embedding", cfg.NODE_EMB.OUTPUT_SIZES, x, cfg.NODE_EMB.BN, cfg.NODE_EMB.END_WITH_RELU)
    g = mlp_forward(sd, "edge_embedding", cfg.EDGE_EMB.OUTPUT_SIZES, edge_attr, cfg.EDGE_EMB.BN, cfg.EDGE_EMB.END_WITH_RELU)
    h0, g0 = h, g
    preds_edge, preds_node, preds_class = [], [], []
    if cfg.AGGR_TYPE == "per_type":
        num_types = {"not": cfg.NUM_JOINTS, "per_body_part": 6, "left_right": 9}[cfg.NODE_TYPE_SUMMARY]

    def heads_node(hh):
        return (mlp_forward(sd, "node_classification", cfg.NODE_CLASS.OUTPUT_SIZES, hh, cfg.BN).squeeze(),
                mlp_forward(sd, "classification", cfg.CLASS.OUTPUT_SIZES, hh, cfg.BN))

    for i in range(cfg.STEPS):                                    # :75
        if cfg.SKIP:                                              # :76-78
            h = np.concatenate([h0, h], 1)
            g = np.concatenate([g0, g], 1)
        if cfg.AGGR_TYPE == "agnostic":
            h, g = mp_layer(sd, "mpn_node_cls", h, g, edge_index, cfg.AGGR, cfg.USE_NODE_UPDATE_MLP)
        else:
            h, g = type_aware_layer(sd, "mpn_node_cls", h, g, edge_index, node_types, cfg.AGGR,
                                    cfg.AGGR_SUB, num_types, getattr(cfg, "UPDATE_TYPE", "mlp"))
        if i >= cfg.STEPS - cfg.AUX_LOSS_STEPS - 1:               # :81-84
            pn, pc = heads_node(h)
            preds_node.append(pn)
            preds_class.append(pc)
            preds_edge.append(mlp_forward(sd, "edge_classification", cfg.EDGE_CLASS.OUTPUT_SIZES, g, cfg.BN).squeeze())
    assert cfg.NODE_STEPS == 0                                    # :87-91 unreachable for per_type
    pn, pc = heads_node(h)                                        # :93-94
    preds_node.append(pn)
    preds_class.append(pc)
    return preds_edge, preds_node, preds_class
