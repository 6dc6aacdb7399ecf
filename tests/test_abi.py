"""CPU checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
include/pgmp.h declares; the host mirrors refuse CPU tensors (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

import pgmp_b200
import pgmp_b200._native as nv
from pgmp_b200.graph_constructor import get_graph_constructor
from pgmp_b200.Models.MessagePassingNetwork import get_mpn_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "pgmp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(pgmp_[a-z_]+)\s*\(", text))


def test_library_exports_every_declared_symbol():
    lib = nv.lib()
    declared = _declared_functions()
    assert declared == set(nv.SYMBOLS), declared ^ set(nv.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.pgmp_version() == 100
    assert isinstance(lib.pgmp_kernel_launches(), int)


def test_struct_sizes_match_header():
    """sizeof() of every ctypes mirror equals the C compiler's (catches field drift)."""
    import subprocess
    import tempfile
    src = '#include <stdio.h>\n#include "pgmp.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(pgmp_gc_params),' \
          ' sizeof(pgmp_gc_outputs), sizeof(pgmp_mlp), sizeof(pgmp_mpn_params), sizeof(pgmp_group_params),' \
          ' sizeof(pgmp_gc_assembly));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(v) for v in subprocess.check_output([exe]).split()]
    mirrors = [nv.GcParams, nv.GcOutputs, nv.Mlp, nv.MpnParams, nv.GroupParams, nv.GcAssembly]
    assert sizes == [ctypes.sizeof(m) for m in mirrors]


def test_workspace_queries_and_error_text_without_gpu():
    lib = nv.lib()
    p = nv.GcParams(batch=2, num_joints=17, height=128, width=128, pool_kernel=5, top_k=10, use_threshold=1,
                    threshold=1.0, graph_type=0, knn_k=50, edge_features=3, norm_factor=128.0, cand_capacity=4096,
                    max_det_per_type=256, max_nodes=2048)
    assert lib.pgmp_gc_workspace_bytes(p) > 2 * 2048 * 64 * 4
    # invalid arguments are rejected before any CUDA call, with a message
    p.pool_kernel = 4
    rc = lib.pgmp_gc_detect(p, None, None)
    assert rc == -1 and b"pool_kernel" in lib.pgmp_last_error()


def test_host_mirrors_refuse_cpu_tensors():
    cfg = pgmp_b200.config.bench_gc_config(k=5)
    z = torch.zeros(1, 17, 32, 32)
    with pytest.raises(RuntimeError, match="CUDA"):
        get_graph_constructor(cfg, scoremaps=z, tagmaps=z, features=torch.zeros(1, 8, 32, 32), joints_gt=None,
                              factor_list=None, masks=None, device="cpu", testing=True, heatmaps=None, num_joints=17)
    model = get_mpn_model(pgmp_b200.config.flagship_mpn_config()).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        model(torch.zeros(4, 128), torch.zeros(2, 19), torch.zeros(2, 2, dtype=torch.long),
              node_types=torch.zeros(4, dtype=torch.long))


def test_factories_follow_the_reference_contract():
    with pytest.raises(NotImplementedError):
        get_mpn_model(pgmp_b200.config.default_mpn_config(NAME="VanillaMPN"))
    m = get_mpn_model(pgmp_b200.config.flagship_mpn_config())
    names = dict(m.named_parameters())
    # SURVEY.md 8b state-dict contract (flagship: 368 596 parameters)
    assert sum(p.numel() for p in m.parameters()) == 368596
    assert names["mpn_node_cls.mlp_edge.0.weight"].shape == (64, 384)
    assert names["mpn_node_cls.mlp_node.mlp.16.0.weight"].shape == (64, 192)
    assert names["mpn_node_cls.update_mlp.0.weight"].shape == (64, 1088)
    assert names["mpn_node_cls.attn_net.0.weight"].shape == (1, 64)
    assert names["edge_embedding.9.weight"].shape == (64, 64) and "edge_embedding.8.weight" in names
    a = get_mpn_model(pgmp_b200.config.agnostic_mpn_config())
    assert sum(p.numel() for p in a.parameters()) == 101203
    with pytest.raises(RuntimeError, match="no CPU path"):      # per-type training runs in libpgmp.so as well
        m.train()(torch.zeros(1), torch.zeros(1), torch.zeros(1), node_types=torch.zeros(1, dtype=torch.int64))


def test_training_mlp_descriptor_follows_make_mlp():
    """Host logic of the training path (no GPU): a ``make_mlp`` chain maps onto ``pgmp_mlp_train`` layer by layer --
    ReLU before BatchNorm (layers.py:11-14), last Linear bare unless END_WITH_RELU, offsets from the flat layout."""
    import torch

    import pgmp_b200
    from pgmp_b200.Models.MessagePassingNetwork import _mlp_train_struct, get_mpn_model

    model = get_mpn_model(pgmp_b200.config.agnostic_mpn_config(17, STEPS=2))
    offsets, off = {}, 0
    for n, p in model.named_parameters():
        offsets[n] = off
        off += (p.numel() + 3) // 4 * 4
    m = _mlp_train_struct(model.edge_embedding, "edge_embedding", offsets)
    assert m.n_layers == 4 and list(m.dims)[:5] == [19, 32, 64, 64, 64]
    assert list(m.relu)[:4] == [1, 1, 1, 0] and list(m.bn)[:4] == [1, 1, 1, 0]
    assert m.w[1] == offsets["edge_embedding.3.weight"] and m.gamma[1] == offsets["edge_embedding.5.weight"]
    assert m.running_mean[0] == model.edge_embedding[2].running_mean.data_ptr()
    h = _mlp_train_struct(model.classification, "classification", offsets)
    assert h.n_layers == 3 and list(h.dims)[:4] == [64, 64, 32, 17] and list(h.relu)[:3] == [1, 1, 0] and not any(h.bn)
    assert all(o % 4 == 0 for o in offsets.values())            # 16-byte aligned pieces (vectorised loads)
    cfg = pgmp_b200.config.agnostic_mpn_config(17, STEPS=2)
    cfg.NODE_EMB.END_WITH_RELU = True
    e = _mlp_train_struct(get_mpn_model(cfg).node_embedding, "node_embedding",
                          {n: 0 for n, _ in get_mpn_model(cfg).named_parameters()})
    assert list(e.relu)[:3] == [1, 1, 1] and list(e.bn)[:3] == [1, 1, 1]


def test_training_mode_rejects_cpu_tensors():
    import pytest
    import torch

    import pgmp_b200
    from pgmp_b200.Models.MessagePassingNetwork import get_mpn_model

    model = get_mpn_model(pgmp_b200.config.agnostic_mpn_config(17, STEPS=2)).train()
    with pytest.raises(RuntimeError, match="no CPU path"):
        model(torch.zeros(4, 128), torch.zeros(2, 19), torch.zeros(2, 2, dtype=torch.int64), node_types=torch.zeros(4, dtype=torch.int64))
    flagship = get_mpn_model(pgmp_b200.config.flagship_mpn_config(17, STEPS=2)).train()
    with pytest.raises(RuntimeError, match="no CPU path"):
        flagship(torch.zeros(4, 128), torch.zeros(2, 19), torch.zeros(2, 2, dtype=torch.int64), node_types=torch.zeros(4, dtype=torch.int64))
    hier = get_mpn_model(pgmp_b200.config.flagship_mpn_config(17, STEPS=2, UPDATE_TYPE="hierarch_mlp")).train()
    with pytest.raises(NotImplementedError):     # the training kernels cover UPDATE_TYPE mlp
        hier(torch.zeros(4, 128), torch.zeros(2, 19), torch.zeros(2, 2, dtype=torch.int64), node_types=torch.zeros(4, dtype=torch.int64))
