"""Graph-constructor oracle (numpy).  TEST INFRASTRUCTURE -- see oracle/__init__.py.

Follows ``src/graph_constructor/ConstructGraph.py`` (``CG.py``) and
``src/Utils/Utils.py`` of the reference; inference branch only (no labels).
All index outputs are int64, all float outputs float32, like the reference.
"""

import numpy as np

KNN_K = 50           # CG.py:365 (hard-coded k)
NO_THRESHOLD_K = 20  # CG.py:1185


# --------------------------------------------------------------------------- NMS
def non_maximum_suppression(scoremap, pool_kernel):
    """Utils.py:15-20.  ``MaxPool2d(k, 1, k//2)`` (-inf padding) ``== x`` as float.

    scoremap: [J,H,W] float32.  The reference's ``threshold`` argument is unused
    (Utils.py:15) and therefore absent here.
    """
    assert pool_kernel % 2 == 1                                    # Utils.py:16
    scoremap = np.asarray(scoremap, dtype=np.float32)
    J, H, W = scoremap.shape
    r = pool_kernel // 2
    pad = np.full((J, H + 2 * r, W + 2 * r), -np.inf, dtype=np.float32)
    pad[:, r:r + H, r:r + W] = scoremap
    # separable max: rows then columns
    hmax = pad[:, :, 0:W].copy()
    for d in range(1, pool_kernel):
        np.maximum(hmax, pad[:, :, d:d + W], out=hmax)
    pooled = hmax[:, 0:H, :].copy()
    for d in range(1, pool_kernel):
        np.maximum(pooled, hmax[:, d:d + H, :], out=pooled)
    return (pooled == scoremap).astype(np.float32)


def _topk_desc_stable(vals, k):
    """Indices of the k largest entries, ordered (value desc, index asc).

    ``torch.topk`` (CG.py:1170,1187) leaves the order among equal values
    unspecified; the build's convention is lowest flat index first.
    """
    n = vals.shape[0]
    k = min(k, n)
    if k == 0:
        return np.zeros(0, dtype=np.int64)
    kth = np.partition(vals, n - k)[n - k]
    cand = np.flatnonzero(vals >= kth)
    order = np.argsort(-vals[cand], kind="stable")
    return cand[order[:k]].astype(np.int64)


def joint_det_from_scoremap(scoremap, num_joints, threshold, pool_kernel, mask=None, hybrid_k=5):
    """CG.py:1161-1196.

    threshold is None -> no-threshold path (top-20 per type, scores + 1e-10).
    Returns (joint_det [N,3] int64 (x,y,type), joint_scores [N] float32).
    """
    scoremap = np.asarray(scoremap, dtype=np.float32)
    J, H, W = scoremap.shape
    assert J == num_joints
    joint_map = non_maximum_suppression(scoremap, pool_kernel)      # CG.py:1162
    if mask is not None:
        joint_map = joint_map * np.asarray(mask, dtype=np.float32)[None]   # CG.py:1163-1164
    s = scoremap * joint_map                                        # CG.py:1165
    flat = s.reshape(J, -1)
    if threshold is not None:
        k = hybrid_k
        container = np.zeros_like(flat)                             # CG.py:1171
        for j in range(J):
            idx = _topk_desc_stable(flat[j], k)                     # CG.py:1170
            container[j, idx] = flat[j, idx]                        # CG.py:1172
        t1, y1, x1 = np.nonzero(container.reshape(J, H, W))         # CG.py:1174 (type,y,x) order
        s_thr = np.where(s < np.float32(threshold), np.float32(0), s)   # CG.py:1177
        t2, y2, x2 = np.nonzero(s_thr)                              # CG.py:1178
        top = np.stack([x1, y1, t1], 1).astype(np.int64)            # CG.py:1180
        thr = np.stack([x2, y2, t2], 1).astype(np.int64)            # CG.py:1181
        det = cat_unique(top, thr)                                  # CG.py:1182
        scores = s[det[:, 2], det[:, 1], det[:, 0]].astype(np.float32)   # CG.py:1183
        return det, scores
    k = NO_THRESHOLD_K
    container = np.zeros_like(flat)
    for j in range(J):
        idx = _topk_desc_stable(flat[j], k)                         # CG.py:1187
        container[j, idx] = flat[j, idx] + np.float32(1e-10)        # CG.py:1189
    t, y, x = np.nonzero(container.reshape(J, H, W))                # CG.py:1191
    scores = container.reshape(J, H, W)[t, y, x].astype(np.float32)  # CG.py:1192
    assert len(t) == k * num_joints                                 # CG.py:1193
    det = np.stack([x, y, t], 1).astype(np.int64)                   # CG.py:1195
    return det, scores


def cat_unique(t1, t2):
    """CG.py:1199-1209: t1 followed by the rows of t2 that do not occur in t1."""
    assert t1.ndim == 2 and t2.ndim == 2
    if len(t1) == 0 or len(t2) == 0:
        return np.concatenate([t1, t2], 0)
    # rows are (x, y, type) with small non-negative ints -> hashable scalar key
    def key(t):
        return (t[:, 2].astype(np.int64) << 40) | (t[:, 1].astype(np.int64) << 20) | t[:, 0].astype(np.int64)
    keep = ~np.isin(key(t2), key(t1))
    return np.concatenate([t1, t2[keep]], 0)


# ------------------------------------------------------------------- edge index
def _to_undirected_no_self_loops(src, dst, n):
    """``to_undirected`` (cat + flip, coalesce = sort unique by src*N+dst) then
    ``remove_self_loops`` -- CG.py:366-367 / 379-380."""
    s = np.concatenate([src, dst]).astype(np.int64)
    d = np.concatenate([dst, src]).astype(np.int64)
    key = np.unique(s * n + d)
    s, d = key // n, key % n
    keep = s != d
    return np.stack([s[keep], d[keep]], 0).astype(np.int64)


def knn_mpn_graph(joint_det, k=KNN_K):
    """CG.py:363-368.  kNN on float32 (x,y); edge = [neighbour -> node]; symmetrised.

    Tie rule (PARITY UNPINNED against torch_cluster, see oracle/__init__.py):
    the k nearest *other* nodes ordered by (squared distance asc, index asc).
    torch_cluster asks for k+1 neighbours including the node itself and drops
    row == col; with a strict-less insertion scan that is the same set as long
    as the node is among its own k+1 nearest, which always holds (d = 0).
    Squared distances of integer pixel coordinates are exact in float32.
    """
    n = len(joint_det)
    if n <= 1:
        return np.zeros((2, 0), dtype=np.int64)
    pos = joint_det[:, :2].astype(np.int64)
    d2 = ((pos[:, None, :] - pos[None, :, :]) ** 2).sum(-1)        # [query, cand]
    big = np.iinfo(np.int64).max
    d2[np.arange(n), np.arange(n)] = big                           # never pick self
    kk = min(k, n - 1)
    order = np.argsort(d2, axis=1, kind="stable")[:, :kk]          # (d2 asc, idx asc)
    query = np.repeat(np.arange(n), kk)
    nbr = order.reshape(-1)
    return _to_undirected_no_self_loops(nbr, query, n)


def fully_connected_mpn_graph(joint_det):
    """CG.py:376-381: all ordered pairs i != j, sorted by (src, dst)."""
    n = len(joint_det)
    src = np.repeat(np.arange(n), n)
    dst = np.tile(np.arange(n), n)
    keep = src != dst
    return np.stack([src[keep], dst[keep]], 0).astype(np.int64)


# --------------------------------------------------------------- node/edge features
def construct_mpn_graph(joint_det, features, graph_type, num_joints, norm_factor,
                        edge_features_to_use=("position", "connection_type")):
    """CG.py:251-361 (the feature sets in scope: position / connection_type / angle).

    features [C,H,W]; returns x [N,C] (dtype of features), edge_attr [E,F] float32,
    edge_index [2,E] int64.  ``norm_factor`` = max(H,W) of the scoremap if
    NORM_NODE_DISTANCE else 1 (CG.py:311-314).
    """
    assert len(edge_features_to_use) >= 1                           # CG.py:260
    jx, jy, jt = joint_det[:, 0], joint_det[:, 1], joint_det[:, 2]
    x = np.ascontiguousarray(features[:, jy, jx].T)                 # CG.py:265,269
    if graph_type == "fully":
        edge_index = fully_connected_mpn_graph(joint_det)           # CG.py:272
    elif graph_type == "knn":
        edge_index = knn_mpn_graph(joint_det)                       # CG.py:274
    else:
        raise NotImplementedError(graph_type)
    src, dst = edge_index
    E = edge_index.shape[1]
    two_hot = np.zeros((E, num_joints), dtype=np.float32)           # CG.py:305
    two_hot[np.arange(E), jt[src]] = 1                              # CG.py:306
    two_hot[np.arange(E), jt[dst]] = 1                              # CG.py:307
    nf = np.float32(norm_factor)
    ea_y = (jy[dst] - jy[src]).astype(np.float32) / nf              # CG.py:316
    ea_x = (jx[dst] - jx[src]).astype(np.float32) / nf              # CG.py:317
    feats = set(edge_features_to_use)
    if feats == {"position", "connection_type"}:                    # CG.py:323-325
        edge_attr = np.concatenate([ea_x[:, None], ea_y[:, None], two_hot], 1)
    elif feats == {"connection_type"}:                              # CG.py:326-328
        edge_attr = two_hot
    elif feats == {"position"}:                                     # CG.py:331-332
        edge_attr = np.stack([ea_x, ea_y], 1)
    elif feats == {"position", "angle", "connection_type"}:         # CG.py:319-321,333-335
        a_x = (jx[src] - jx[dst]).astype(np.float32)
        a_y = (jy[src] - jy[dst]).astype(np.float32)
        with np.errstate(all="ignore"):
            theta = np.abs(np.arccos(a_x * (np.float32(1) / np.sqrt(a_x ** 2 + a_y ** 2)))).astype(np.float32)
        theta[np.isnan(theta)] = 0.0
        edge_attr = np.concatenate([ea_x[:, None], ea_y[:, None], theta[:, None], two_hot], 1)
    else:
        raise NotImplementedError(sorted(feats))
    return x, edge_attr.astype(np.float32), edge_index


def construct_graph(scoremaps, tagmaps, features, cfg, num_joints, masks=None):
    """CG.py:46-68, 100-103, 206-249 (``joints_gt is None``).

    scoremaps [B,J,H,W], tagmaps [B,J,H,W] or [B,J,H,W,T], features [B,C,H,W].
    cfg: object with the GC attributes the reference reads (CG.py:23-44).
    Returns a dict with the non-None slots of the reference's 15-tuple.
    """
    B = scoremaps.shape[0]
    thr = cfg.DETECT_THRESHOLD if cfg.DETECT_THRESHOLD <= 1.5 else None    # CG.py:28
    norm = max(scoremaps.shape[3], scoremaps.shape[2]) if cfg.NORM_NODE_DISTANCE else 1
    xs, eas, eis, dets, scs, tags, bidx = [], [], [], [], [], [], []
    n_nodes = [0]
    for b in range(B):                                              # CG.py:58
        mask = masks[b] if cfg.MASK_CROWDS else None                # CG.py:59-68
        det, sc = joint_det_from_scoremap(scoremaps[b], num_joints, thr, cfg.POOL_KERNEL_SIZE,
                                          mask=mask, hybrid_k=cfg.HYBRID_K)
        x, ea, ei = construct_mpn_graph(det, features[b], cfg.GRAPH_TYPE, num_joints, norm,
                                        cfg.EDGE_FEATURES_TO_USE)   # CG.py:100-102
        tg = tagmaps[b, det[:, 2], det[:, 1], det[:, 0]]            # CG.py:103
        xs.append(x); eas.append(ea); eis.append(ei + n_nodes[-1])  # CG.py:222-223
        dets.append(det); scs.append(sc); tags.append(tg)
        bidx.append(np.full(len(det), b, dtype=np.int64))           # CG.py:207
        n_nodes.append(n_nodes[-1] + len(det))
    return dict(
        x=np.concatenate(xs, 0), edge_attr=np.concatenate(eas, 0),
        edge_index=np.concatenate(eis, 1), joint_det=np.concatenate(dets, 0),
        joint_scores=np.concatenate(scs, 0), batch_index=np.concatenate(bidx, 0),
        joint_tags=np.concatenate(tags, 0), num_nodes=np.diff(n_nodes),
        num_edges=np.array([e.shape[1] for e in eis], dtype=np.int64),
    )
