"""Heads -> threshold -> multicut (GAEC) -> person grouping oracle (numpy).
TEST INFRASTRUCTURE -- see oracle/__init__.py.

Follows ``src/valid.py:109-122``, ``src/Utils/Utils.py`` (pred_to_ann :1445-1457,
pred_to_person :499-514, graph_cluster_to_persons :672-743, subgraph_mask :981-993)
and ``src/Utils/correlation_clustering/correlation_clustering_utils.py``.

GAEC itself lives in the reference's missing native module
``andres_graph_wrapper`` (no source, no version pin) -> PARITY UNPINNED for the
solver; the restatement follows the published greedy additive edge contraction
(Keuper et al. 2015; andres/graph ``multicut/greedy-additive.hxx``) with this
tie rule: among live edges of equal (maximal) weight the one with the smallest
(a, b), a < b, is contracted first, where a cluster is named by its smallest
member.  Weights are accumulated in float64.
"""

import heapq

import numpy as np


def sigmoid(v):
    v = np.asarray(v, dtype=np.float32)
    return (np.float32(1) / (np.float32(1) + np.exp(-v))).astype(np.float32)


def softmax(v):
    v = np.asarray(v, dtype=np.float32)
    e = np.exp(v - v.max(axis=1, keepdims=True))
    return (e / e.sum(axis=1, keepdims=True)).astype(np.float32)


def threshold_subgraph(node_prob, edge_index, edge_prob, th):
    """Utils.py:1450-1451 (``torch_geometric.utils.subgraph`` without relabelling)."""
    keep = np.asarray(node_prob) > np.float32(th)
    src, dst = edge_index
    m = keep[src] & keep[dst]                                     # Utils.py:981-993
    return edge_index[:, m], edge_prob[m], keep


def multicut_weights(edge_index, edge_prob):
    """extract_edge_matrix(update=True) + cluster_andres_graph(complete=False),
    correlation_clustering_utils.py:99-136, 209-227.

    Returns (a, b, w): undirected edges a < b in edge-list order and
    w = (p_ab + p_ba) / 2 - 0.5 in float32 (p_ba = 0 if the reverse edge is
    absent; if *no* lower-triangle entry exists at all the matrix is mirrored
    instead of averaged, :117-121).
    """
    src, dst = edge_index
    if len(src) == 0:
        z = np.zeros(0, dtype=np.int64)
        return z, z, np.zeros(0, dtype=np.float32)
    n = int(edge_index.max()) + 1                                 # to_dense_adj sizing, :112
    A = np.zeros((n, n), dtype=np.float32)
    np.add.at(A, (src, dst), edge_prob.astype(np.float32))        # to_dense_adj scatters with add
    if np.tril(A).sum() == 0:                                     # :114-117
        A = A + A.T
    else:
        A = (A + A.T) / np.float32(2)                             # :118-121
    up = src < dst                                                # :224-225
    a, b = src[up], dst[up]
    w = A[a, b] - np.float32(0.5)                                 # :221
    return a.astype(np.int64), b.astype(np.int64), w.astype(np.float32)


def gaec(a, b, w, n):
    """Greedy additive edge contraction on the undirected weighted graph
    (a[i], b[i], w[i]), a < b, n vertices.  Returns the cluster representative
    (smallest member) of every vertex.  See the module docstring for the tie rule.
    """
    adj = [dict() for _ in range(n)]
    for ai, bi, wi in zip(a.tolist(), b.tolist(), w.astype(np.float64).tolist()):
        adj[ai][bi] = adj[ai].get(bi, 0.0) + wi
        adj[bi][ai] = adj[bi].get(ai, 0.0) + wi
    heap = [(-wt, u, v) for u in range(n) for v, wt in adj[u].items() if u < v]
    heapq.heapify(heap)
    rep = np.arange(n)
    alive = np.ones(n, dtype=bool)
    while heap:
        negw, u, v = heapq.heappop(heap)
        if not (alive[u] and alive[v]) or adj[u].get(v) != -negw:
            continue                                              # stale entry (edition check)
        if -negw < 0.0:
            break                                                 # "there must be negative weights", :213,222
        # contract v into u (u < v: the cluster keeps its smallest member as name)
        alive[v] = False
        rep[rep == v] = u
        del adj[u][v]
        for p, wt in adj[v].items():
            if p == u:
                continue
            del adj[p][v]
            nw = adj[u].get(p, 0.0) + wt if p in adj[u] else wt
            adj[u][p] = nw
            adj[p][u] = nw
            heapq.heappush(heap, (-nw, min(u, p), max(u, p)))
        adj[v] = {}
    return rep


def connected_component_labels(rep):
    """scipy ``connected_components`` numbering: components in order of their
    smallest node index (Utils.py:688-691).  ``rep`` names each node's cluster."""
    labels = np.full(len(rep), -1, dtype=np.int64)
    nxt = 0
    seen = {}
    for i, r in enumerate(rep.tolist()):
        if r not in seen:
            seen[r] = nxt
            nxt += 1
        labels[i] = seen[r]
    return labels


def graph_cluster_to_persons(joint_det, node_prob, person_labels, class_prob, num_joints):
    """Utils.py:672-743 (``scores_for_poses=None``, ``allow_single_joint_persons=False``).

    Returns (persons [P,J,3] float64, mutant_detected)."""
    persons, mutant = [], False
    n_comp = int(person_labels.max()) + 1 if len(person_labels) else 0
    for c in range(n_comp):
        sel = person_labels == c
        pj = joint_det[sel].copy()
        ps = node_prob[sel]
        if class_prob is not None:
            pj[:, 2] = np.argmax(class_prob[sel], axis=1)         # :699-702
        if len(pj) > num_joints:
            mutant = True                                         # :703-706
        if len(pj) > 1:                                           # :708
            kp = np.zeros([num_joints, 3])
            for t in range(num_joints):
                s = pj[:, 2] == t
                if s.any():
                    k = int(np.argmax(ps[s]))                     # :718
                    kp[t] = pj[s][k]                              # :719
                    kp[t, 2] = ps[s].max()                        # :720
            if (kp[:, 2] > 0).sum() > 0:                          # :725
                persons.append(kp)
    return np.array(persons), mutant


def threshold_clusters(edge_index, edge_prob, n, edge_threshold=0.8):
    """``CC_METHOD == "threshold"`` (Utils.py:508-509): the kept edges with probability > 0.8 are the solution; their
    connected components are the persons.  Returns each node's representative = smallest node of its component."""
    rep = np.arange(n, dtype=np.int64)

    def find(i):
        while rep[i] != i:
            rep[i] = rep[rep[i]]
            i = rep[i]
        return i
    for s, d in edge_index[:, edge_prob > np.float32(edge_threshold)].T.tolist():
        a, b = find(s), find(d)
        if a != b:
            rep[max(a, b)] = min(a, b)
    return np.array([find(i) for i in range(n)], dtype=np.int64)


def greedy_person_construction(joint_det, node_prob, edge_index, edge_prob, class_prob, num_joints):
    """``CC_METHOD == "greedy"`` (Utils.py:517-626).  Returns (persons [P,J,3] float64, taken [N] int64: the core node
    that claimed each node, -1 for unclaimed nodes)."""
    jd = np.array(joint_det, dtype=np.int64, copy=True)
    n = len(jd)
    if class_prob is not None:
        jd[:, 2] = np.argmax(class_prob, axis=1)                  # :528-529
    adj = np.zeros((n, n), dtype=np.float64)                      # :530
    adj[edge_index[0], edge_index[1]] = edge_prob                 # :531
    adj = (adj.T + adj) / 2.0                                     # :532
    np.fill_diagonal(adj, 1.0)                                    # :533
    taken = np.full(n, -1, dtype=np.int64)
    by_type = [np.flatnonzero(jd[:, 2] == t) for t in range(num_joints)]
    for t in range(num_joints):                                   # :549-583
        for i in by_type[t].tolist():
            if taken[i] != -1 or node_prob[i] < 0.5:
                continue
            taken[i] = i
            for j in range(num_joints):
                if j == t or len(by_type[j]) == 0:
                    continue
                row = adj[i, by_type[j]]
                k = int(np.argmax(row))                           # first maximum; entries of other types count as 0
                score, target = row[k], int(by_type[j][k])
                if score == 0.0:
                    continue
                owner = taken[target]
                if owner != -1 and adj[owner, target] > score:
                    continue
                taken[target] = i
    persons = []
    for c in range(int(taken.max()) + 1 if n else 0):             # :586-613
        sel = taken == c
        if sel.sum() > 1:
            kp = np.zeros([num_joints, 3])
            for t in range(num_joints):
                s = sel & (jd[:, 2] == t)
                if s.any():
                    k = np.flatnonzero(s)[int(np.argmax(node_prob[s]))]
                    kp[t] = jd[k]
                    kp[t, 2] = node_prob[s].max()
            if (kp[:, 2] > 0).sum() > 0:
                persons.append(kp)
    return np.array(persons), taken


def pred_to_person(joint_det, node_logits, edge_index, edge_logits, class_logits, node_threshold, num_joints,
                   cc_method="GAEC"):
    """valid.py:109-111 + Utils.py:1448-1457 + :499-514 for ``CC_METHOD in ("GAEC", "threshold", "greedy")``.

    Returns (persons, mutants, person_labels) or None when the reference's
    ``pred_to_ann`` returns None before grouping (no edge survives)."""
    p_node = sigmoid(node_logits)
    p_edge = sigmoid(edge_logits)
    p_cls = softmax(class_logits) if class_logits is not None else None
    ei, pe, _ = threshold_subgraph(p_node, edge_index, p_edge, node_threshold)
    if ei.shape[1] == 0:
        return None                                               # Utils.py:1452,1457
    n = len(joint_det)
    if cc_method == "greedy":
        persons, taken = greedy_person_construction(np.asarray(joint_det), p_node, ei, pe, p_cls, num_joints)
        return persons, False, taken                              # Utils.py:505-507
    if cc_method == "threshold":
        rep = threshold_clusters(ei, pe, n)
    else:
        a, b, w = multicut_weights(ei, pe)
        rep = gaec(a, b, w, n)
    labels = connected_component_labels(rep)
    persons, mutant = graph_cluster_to_persons(np.asarray(joint_det), p_node, labels, p_cls, num_joints)
    return persons, mutant, labels
