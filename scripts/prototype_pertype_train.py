"""Design prototype (numpy, float64) of the per-type training kernels planned in DESIGN.md section 7a -- NOT product code and
NOT the oracle.  It states, operation by operation, what the CUDA version of one ``TypeAwareMPNLayer`` training step will do
(type-sorted edge order, per-node tables, per-(target, type) attention bins, per-node sums before the weight products in the
reverse pass) so that the algebra is settled before any kernel is written; ``tests/test_pertype_train_plan.py`` checks it
against the tape of ``oracle/mpn_train.py`` (which is pinned to the reference under float64 autograd).

Reference: ``TypeAwareMPNLayer.forward`` / ``TypeAwareNodeUpdate`` (src/Models/MessagePassingNetwork/layers.py:157-274),
AGGR_SUB ``node_edge_attn`` (one attention logit per edge) or ``node_edge_attn_per_type`` (column = source type).
"""
import numpy as np

D = 64
EPS = 1e-12


def relu(v):
    return np.maximum(v, 0.0)


def layer_forward_backward(P, x, e, src, dst, node_types, num_types, per_type_attn, d_h, d_g):
    """One layer: x [N, nd], e [E, ed] -> h' [N, 64], g' [E, 64]; then the reverse pass for upstream gradients
    d_h (of h') and d_g (of g').  ``P`` maps parameter names (without the layer prefix) to float64 arrays.
    Returns (h', g', dx, de, {name: gradient})."""
    N, nd = x.shape
    E, ed = e.shape
    T = num_types
    W1, b1 = P["mlp_edge.0.weight"], P["mlp_edge.0.bias"]
    W2, b2 = P["mlp_edge.2.weight"], P["mlp_edge.2.bias"]
    Wu, bu = P["update_mlp.0.weight"], P["update_mlp.0.bias"]
    wa, ba = P["attn_net.0.weight"], P["attn_net.0.bias"]
    Wm = np.stack([P["mlp_node.mlp.%d.0.weight" % t] for t in range(17)])          # [17, 64, nd + 64]
    bm = np.stack([P["mlp_node.mlp.%d.0.bias" % t] for t in range(17)])

    # ---- type-sorted edge order: everything per edge is order-agnostic, only the outputs are un-permuted
    etype = node_types[src]
    perm = np.argsort(etype, kind="stable")
    s_src, s_dst, s_type, s_e = src[perm], dst[perm], etype[perm], e[perm]
    gstart = np.searchsorted(s_type, np.arange(T + 1))                              # 18-entry group table

    # ---- forward
    tabP, tabQ = x @ W1[:, :nd].T, x @ W1[:, nd:2 * nd].T                           # per node
    hid = relu(s_e @ W1[:, 2 * nd:].T + b1 + tabP[s_dst] + tabQ[s_src])
    g1 = relu(hid @ W2.T + b2)
    tabR = np.einsum("nk,tok->nto", x, Wm[:, :, :nd]) + bm[None]                    # [N, 17, 64]: one block-strided product
    m = np.zeros((E, D))
    for t in range(T):                                                              # one weight matrix per tile group
        a, b = gstart[t], gstart[t + 1]
        m[a:b] = relu(g1[a:b] @ Wm[t, :, nd:].T + tabR[s_dst[a:b], t])
    logits = g1 @ wa.T + ba                                                         # [E, 1] or [E, 17]
    a_e = logits[np.arange(E), s_type] if per_type_attn else logits[:, 0]
    key = s_type * N + s_dst                                                        # (type, target) bins
    order = np.argsort(key, kind="stable")
    bins = np.flatnonzero(np.r_[True, key[order][1:] != key[order][:-1], True])
    alpha = np.zeros(E)
    U = np.zeros((N, T, D))
    for i in range(len(bins) - 1):                                                  # one warp per bin
        rows = order[bins[i]:bins[i + 1]]
        ex = np.exp(a_e[rows] - a_e[rows].max())
        alpha[rows] = ex / (ex.sum() + EPS)
        U[s_dst[rows[0]], s_type[rows[0]]] = (alpha[rows, None] * m[rows]).sum(0)
    h1 = relu(U.reshape(N, T * D) @ Wu.T + bu)
    g_out = np.empty_like(g1)
    g_out[perm] = g1

    # ---- reverse pass
    G = {}
    dpre = d_h * (h1 > 0)
    G["update_mlp.0.weight"], G["update_mlp.0.bias"] = dpre.T @ U.reshape(N, T * D), dpre.sum(0)
    dU = (dpre @ Wu).reshape(N, T, D)
    dm = np.zeros((E, D))
    da = np.zeros(E)
    for i in range(len(bins) - 1):
        rows = order[bins[i]:bins[i + 1]]
        du = dU[s_dst[rows[0]], s_type[rows[0]]]
        dm[rows] = alpha[rows, None] * du
        dalpha = m[rows] @ du
        da[rows] = alpha[rows] * (dalpha - (alpha[rows] * dalpha).sum())
    dm *= m > 0
    dlogits = np.zeros_like(logits)
    if per_type_attn:
        dlogits[np.arange(E), s_type] = da
    else:
        dlogits[:, 0] = da
    G["attn_net.0.weight"], G["attn_net.0.bias"] = dlogits.T @ g1, dlogits.sum(0)
    dg1 = d_g[perm] + dlogits @ wa
    dR = np.zeros((N, T, D))                                                        # per-(target, type) sums of dm
    np.add.at(dR, (s_dst, s_type), dm)
    dWm, dbm = np.zeros_like(Wm), np.zeros_like(bm)
    for t in range(T):
        a, b = gstart[t], gstart[t + 1]
        dWm[t, :, nd:] = dm[a:b].T @ g1[a:b]                                        # fixed row ranges per type group
        dg1[a:b] += dm[a:b] @ Wm[t, :, nd:]
        dWm[t, :, :nd] = dR[:, t].T @ x
        dbm[t] = dR[:, t].sum(0)
    for t in range(17):
        G["mlp_node.mlp.%d.0.weight" % t], G["mlp_node.mlp.%d.0.bias" % t] = dWm[t], dbm[t]
    dx = np.einsum("nto,tok->nk", dR, Wm[:, :, :nd])
    dg1 *= g1 > 0
    G["mlp_edge.2.weight"], G["mlp_edge.2.bias"] = dg1.T @ hid, dg1.sum(0)
    dhid = (dg1 @ W2) * (hid > 0)
    S_dst, S_src = np.zeros((N, D)), np.zeros((N, D))
    np.add.at(S_dst, s_dst, dhid)
    np.add.at(S_src, s_src, dhid)
    dW1 = np.zeros_like(W1)
    dW1[:, :nd], dW1[:, nd:2 * nd], dW1[:, 2 * nd:] = S_dst.T @ x, S_src.T @ x, dhid.T @ s_e
    G["mlp_edge.0.weight"], G["mlp_edge.0.bias"] = dW1, S_dst.sum(0)
    dx += S_dst @ W1[:, :nd] + S_src @ W1[:, nd:2 * nd]
    de = np.empty_like(e)
    de[perm] = dhid @ W1[:, 2 * nd:]
    return h1, g_out, dx, de, G
