"""Import the UNMODIFIED reference hot-path files (TEST / BASELINE INFRASTRUCTURE ONLY).

The files are loaded byte-for-byte from the first root that exists:
``$PGMP_REF_SRC``, ``/root/reference/src`` (this container) or ``baseline/_ref/src``
(the verbatim, git-ignored copy ``oracle/make_ref.py`` makes so that the reference
itself can be timed on the GPU box, where ``/root/reference`` does not exist).
Users: ``tests/golden/make_golden*.py`` (fixtures) and the CPU arms of ``bench.py``
(``--impl reference``, ``cpu_baseline``).  The product never imports it.  The reference's third-party
dependencies that cannot be installed offline are replaced by the small
pure-torch stand-ins below (SURVEY.md 8c lists them); the reference's own files
are loaded byte-for-byte with ``importlib``.

Stand-in semantics that are OUR convention (parity unpinned, see oracle/__init__.py):
``knn_graph`` orders neighbours by (squared distance asc, index asc).
"""

import importlib.util
import inspect
import sys
import types

import torch

import os

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_FILES = [          # the reference files of the hot path, relative to its src/
    "Utils/Utils.py",
    "graph_constructor/ConstructGraph.py",
    "Models/MessagePassingNetwork/utils.py",
    "Models/MessagePassingNetwork/layers.py",
    "Models/MessagePassingNetwork/NodeClassificationMPNSimple.py",
    "Utils/correlation_clustering/correlation_clustering_utils.py",
]


def ref_src():
    """Root of the reference's ``src`` tree, or None when no copy of the reference is reachable."""
    for cand in (os.environ.get("PGMP_REF_SRC"), "/root/reference/src", os.path.join(_ROOT, "baseline", "_ref", "src")):
        if cand and all(os.path.exists(os.path.join(cand, f)) for f in REF_FILES):
            return cand
    return None


class _RefSrc:
    def __format__(self, spec):
        root = ref_src()
        if root is None:
            raise RuntimeError("no copy of the reference is reachable (run oracle/make_ref.py where /root/reference exists)")
        return root


REF_SRC = _RefSrc()


# ----------------------------------------------------------------- torch_scatter
def _scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    assert dim == 0
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    shape = (dim_size,) + tuple(src.shape[1:])
    idx = index.view((-1,) + (1,) * (src.dim() - 1)).expand_as(src)
    if reduce in ("sum", "add"):
        return torch.zeros(shape, dtype=src.dtype).scatter_add_(0, idx, src)
    if reduce == "mean":
        s = torch.zeros(shape, dtype=src.dtype).scatter_add_(0, idx, src)
        c = torch.zeros(dim_size, dtype=src.dtype).scatter_add_(0, index, torch.ones_like(index, dtype=src.dtype))
        return s / c.clamp(min=1).view((-1,) + (1,) * (src.dim() - 1))
    if reduce == "max":
        o = torch.full(shape, float("-inf"), dtype=src.dtype).scatter_reduce_(0, idx, src, "amax", include_self=True)
        return torch.where(torch.isinf(o), torch.zeros_like(o), o)   # torch_scatter fills empty segments with 0
    raise NotImplementedError(reduce)


def _scatter_max(src, index, dim=0, out=None, dim_size=None):
    return _scatter(src, index, dim, None, dim_size, "max"), None


def _scatter_mean(src, index, dim=0, out=None, dim_size=None):
    return _scatter(src, index, dim, None, dim_size, "mean")


def _scatter_softmax(src, index, dim=0, eps=1e-12):
    n = int(index.max()) + 1 if index.numel() else 0
    mx = torch.full((n,), float("-inf"), dtype=src.dtype).scatter_reduce_(0, index, src, "amax", include_self=True)
    ex = (src - mx[index]).exp()
    sm = torch.zeros(n, dtype=src.dtype).scatter_add_(0, index, ex) + eps
    return ex / sm[index]


# ----------------------------------------------------------------- torch_geometric
def _coalesce(edge_index, n):
    key = torch.unique(edge_index[0] * n + edge_index[1])          # sorted
    return torch.stack([key // n, key % n], 0)


def _to_undirected(edge_index, num_nodes=None):
    n = num_nodes if num_nodes is not None else int(edge_index.max()) + 1
    row, col = edge_index
    return _coalesce(torch.stack([torch.cat([row, col]), torch.cat([col, row])], 0), n)


def _remove_self_loops(edge_index, edge_attr=None):
    m = edge_index[0] != edge_index[1]
    return edge_index[:, m], (edge_attr[m] if edge_attr is not None else None)


def _dense_to_sparse(t):
    idx = t.nonzero(as_tuple=False).t().contiguous()
    return idx, t[idx[0], idx[1]]


def _subgraph(subset, edge_index, edge_attr=None, relabel_nodes=False, num_nodes=None):
    assert not relabel_nodes
    m = subset[edge_index[0]] & subset[edge_index[1]]
    return edge_index[:, m], (edge_attr[m] if edge_attr is not None else None)


def _knn_graph(x, k, batch=None, loop=False, flow="source_to_target"):
    n = x.shape[0]
    d2 = ((x[:, None, :] - x[None, :, :]) ** 2).sum(-1)            # float32, exact for pixel coords
    kk = min(k + 1, n)                                             # k+1 incl. self, like torch_cluster
    order = torch.argsort(d2, dim=1, stable=True)[:, :kk]          # (d2 asc, index asc)
    query = torch.arange(n).repeat_interleave(kk)
    nbr = order.reshape(-1)
    keep = query != nbr
    return torch.stack([nbr[keep], query[keep]], 0)                # edge = [neighbour -> node]


def _to_dense_adj(edge_index, batch=None, edge_attr=None):
    n = int(edge_index.max()) + 1
    adj = torch.zeros(1, n, n, dtype=edge_attr.dtype)
    adj[0].index_put_((edge_index[0], edge_index[1]), edge_attr, accumulate=True)
    return adj


class _Data:
    def __init__(self, x=None, edge_index=None, edge_attr=None, **kw):
        self.x, self.edge_index, self.edge_attr = x, edge_index, edge_attr

    @property
    def num_nodes(self):
        return len(self.x)

    def cpu(self):
        return self


class _MessagePassing(torch.nn.Module):
    """PyG >= 1.5 ``MessagePassing`` for flow = source_to_target."""

    def __init__(self, aggr="add", **kw):
        super().__init__()
        self.aggr = aggr

    def _args(self, fn, skip):
        return [p for p in list(inspect.signature(fn).parameters)[skip:]]

    def propagate(self, edge_index, size=None, **kwargs):
        j, i = edge_index[0], edge_index[1]
        n = size[1] if size is not None else None
        pool = dict(kwargs)
        for name in self._args(self.message, 0):
            if name.endswith("_i"):
                pool[name] = kwargs[name[:-2]][i]
            elif name.endswith("_j"):
                pool[name] = kwargs[name[:-2]][j]
        pool.update(index=i, dim_size=n, size=size)
        out = self.message(**{a: pool[a] for a in self._args(self.message, 0)})
        out = self.aggregate(out, **{a: pool[a] for a in self._args(self.aggregate, 1)})
        return self.update(out, **{a: pool[a] for a in self._args(self.update, 1)})

    def message(self, x_j):
        return x_j

    def aggregate(self, inputs, index, dim_size=None):
        return _scatter(inputs, index, 0, None, dim_size, self.aggr)

    def update(self, inputs):
        return inputs


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_shims():
    comp = _module("torch_scatter.composite", scatter_softmax=_scatter_softmax)
    _module("torch_scatter", scatter=_scatter, scatter_max=_scatter_max, scatter_mean=_scatter_mean,
            scatter_softmax=_scatter_softmax, composite=comp)
    gu = _module("torch_geometric.utils", to_undirected=_to_undirected, remove_self_loops=_remove_self_loops,
                 dense_to_sparse=_dense_to_sparse, subgraph=_subgraph, to_dense_adj=_to_dense_adj)
    gn = _module("torch_geometric.nn", knn_graph=_knn_graph, MessagePassing=_MessagePassing)
    gd = _module("torch_geometric.data", Data=_Data)
    _module("torch_geometric", utils=gu, nn=gn, data=gd)


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _lift(path, first, last):
    """Source lines [first, last] (1-based) of a reference file."""
    with open(path) as f:
        return "".join(f.readlines()[first - 1:last])


def load_reference():
    """Returns (ConstructGraph module, NodeClassificationMPNSimple module)."""
    install_shims()
    # Utils/Utils.py cannot be imported whole (matplotlib, tensorboard, missing native lib):
    # exec the two functions the graph constructor needs from their own source lines.
    utils = _module("Utils.Utils")
    _module("Utils", Utils=utils)
    ns = {"torch": torch, "nn": torch.nn}
    exec(_lift(f"{REF_SRC}/Utils/Utils.py", 15, 20), ns)           # non_maximum_suppression
    exec(_lift(f"{REF_SRC}/Utils/Utils.py", 981, 993), ns)         # subgraph_mask
    utils.non_maximum_suppression = ns["non_maximum_suppression"]
    utils.subgraph_mask = ns["subgraph_mask"]
    cg = _load("ref_ConstructGraph", f"{REF_SRC}/graph_constructor/ConstructGraph.py")
    pkg = types.ModuleType("ref_mpn")
    pkg.__path__ = [f"{REF_SRC}/Models/MessagePassingNetwork"]
    sys.modules["ref_mpn"] = pkg
    _load("ref_mpn.utils", f"{REF_SRC}/Models/MessagePassingNetwork/utils.py")
    _load("ref_mpn.layers", f"{REF_SRC}/Models/MessagePassingNetwork/layers.py")
    mpn = _load("ref_mpn.NodeClassificationMPNSimple",
                f"{REF_SRC}/Models/MessagePassingNetwork/NodeClassificationMPNSimple.py")
    return cg, mpn


def load_reference_grouping(gaec_fn):
    """The reference's grouping tail with the missing native GAEC solver replaced by
    ``gaec_fn(a, b, w, n) -> cluster representative per vertex``.  Everything else
    (weight plumbing, dense matrices, connected components, person assembly) is the
    reference's own code: correlation_clustering_utils.py is loaded whole,
    ``pred_to_person`` / ``graph_cluster_to_persons`` are lifted from Utils.py.
    Returns (pred_to_person, subgraph) callables."""
    import numpy as np
    install_shims()
    if not hasattr(np, "int"):
        np.int = int                                               # numpy >= 1.24 dropped the alias used at :239

    class _G:
        def __init__(self, edges, weights, n):
            self.edges, self.weights, self.n = edges, weights, n

    def _cluster_gaec(g):
        rep = gaec_fn(g.edges[0], g.edges[1], g.weights, g.n)
        return (rep[g.edges[0]] != rep[g.edges[1]]).astype(np.int64)   # 1 = cut

    wrapper = _module("Utils.correlation_clustering.andres_graph.andres_graph_wrapper",
                      Graph=_G, cluster_GAEC=_cluster_gaec, cluster_KL=None, cluster_MUT=None)
    _module("Utils.correlation_clustering.andres_graph", andres_graph_wrapper=wrapper)
    if "Utils" not in sys.modules:
        _module("Utils")
    cc_pkg = _module("Utils.correlation_clustering")
    ccu = _load("Utils.correlation_clustering.correlation_clustering_utils",
                f"{REF_SRC}/Utils/correlation_clustering/correlation_clustering_utils.py")
    cc_pkg.correlation_clustering_utils = ccu
    ns = {"torch": torch, "np": np, "cluster_graph": ccu.cluster_graph, "Graph": _Data,
          "dense_to_sparse": _dense_to_sparse}
    exec(_lift(f"{REF_SRC}/Utils/Utils.py", 36, 40), ns)            # to_numpy
    if not hasattr(np, "float"):
        np.float = float                                           # alias used at Utils.py:530, dropped by numpy >= 1.24
    exec(_lift(f"{REF_SRC}/Utils/Utils.py", 517, 626), ns)          # greedy_person_construction (CC_METHOD "greedy")
    exec(_lift(f"{REF_SRC}/Utils/Utils.py", 499, 514), ns)          # pred_to_person
    exec(_lift(f"{REF_SRC}/Utils/Utils.py", 672, 743), ns)          # graph_cluster_to_persons
    return ns["pred_to_person"], _subgraph
