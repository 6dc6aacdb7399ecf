"""ctypes binding of ``libpgmp.so`` (C ABI declared in ``include/pgmp.h``).

The shared library is built in-tree from ``csrc/*.cu`` with
``nvcc -gencode arch=compute_100a,code=sm_100a`` (``build()``; ``__graft_entry__.build()``
calls it).  There is no CPU or PyTorch fallback: if the library cannot be loaded every
entry point raises.
"""

import ctypes as C
import glob
import os
import shutil
import subprocess
import threading

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libpgmp.so")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"] + os.environ.get("PGMP_NVCC_EXTRA", "").split()

MAX_LAYERS = 6
GRAPH_KNN, GRAPH_FULLY = 0, 1
EDGE_FEAT_POSITION, EDGE_FEAT_TYPE = 1, 2
GC_FLAGS = {1: "more NMS maxima than B200_CAND_CAPACITY for some (image, joint)",
            2: "more detections than B200_MAX_DET_PER_TYPE for some (image, joint)",
            4: "more nodes than B200_MAX_NODES in some image",
            8: "no-threshold path: a map has fewer than k pixels (the reference's assert, CG.py:1193)"}
AGGR = {"add": 0, "sum": 0, "max": 1, "mean": 2}
ATTN = {"None": 0, "node_edge_attn": 1, "node_edge_attn_per_type": 2}
PRECISION = {"fp32": 0, "tc": 1}


class GcParams(C.Structure):
    _fields_ = [("batch", C.c_int32), ("num_joints", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
                ("pool_kernel", C.c_int32), ("top_k", C.c_int32), ("use_threshold", C.c_int32),
                ("threshold", C.c_float), ("graph_type", C.c_int32), ("knn_k", C.c_int32),
                ("edge_features", C.c_int32), ("norm_factor", C.c_float), ("cand_capacity", C.c_int32),
                ("max_det_per_type", C.c_int32), ("max_nodes", C.c_int32), ("scoremaps", C.c_void_p),
                ("mask", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_uint64)]


class GcAssembly(C.Structure):
    _fields_ = [("stage1", C.c_void_p * 2), ("stage2", C.c_void_p * 2), ("channels1", C.c_int32), ("h", C.c_int32),
                ("w", C.c_int32), ("mode", C.c_int32), ("n_terms", C.c_int32), ("flip_index", C.c_int32 * 32),
                ("scoremaps_out", C.c_void_p)]


class GcOutputs(C.Structure):
    _fields_ = [("total_nodes", C.c_int64), ("total_edges", C.c_int64), ("features", C.c_void_p),
                ("feat_stride_b", C.c_int64), ("feat_stride_c", C.c_int64), ("feat_stride_y", C.c_int64),
                ("feat_stride_x", C.c_int64), ("channels", C.c_int32), ("tagmaps", C.c_void_p),
                ("tag_dim", C.c_int32), ("x", C.c_void_p), ("edge_attr", C.c_void_p), ("edge_index", C.c_void_p),
                ("joint_det", C.c_void_p), ("joint_scores", C.c_void_p), ("batch_index", C.c_void_p),
                ("joint_tags", C.c_void_p)]


class GatherConvParams(C.Structure):
    _fields_ = [("features", C.c_void_p), ("feat_stride_b", C.c_int64), ("feat_stride_c", C.c_int64),
                ("feat_stride_y", C.c_int64), ("feat_stride_x", C.c_int64), ("cin", C.c_int32), ("height", C.c_int32),
                ("width", C.c_int32), ("cout", C.c_int32), ("out_height", C.c_int32), ("out_width", C.c_int32),
                ("weight_t", C.c_void_p), ("bias", C.c_void_p), ("joint_det", C.c_void_p), ("batch_index", C.c_void_p),
                ("num_nodes", C.c_int64), ("x", C.c_void_p)]


class Mlp(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("dims", C.c_int32 * (MAX_LAYERS + 1)), ("relu", C.c_int32 * MAX_LAYERS),
                ("wt", C.c_void_p * MAX_LAYERS), ("bias", C.c_void_p * MAX_LAYERS), ("post_relu", C.c_int32),
                ("post_scale", C.c_void_p), ("post_shift", C.c_void_p)]


class MpnParams(C.Structure):
    _fields_ = [("num_nodes", C.c_int64), ("num_edges", C.c_int64), ("x", C.c_void_p), ("x_stride_n", C.c_int64),
                ("x_stride_c", C.c_int64), ("edge_attr", C.c_void_p), ("edge_index", C.c_void_p),
                ("node_types", C.c_void_p),
                ("dim", C.c_int32), ("per_type", C.c_int32), ("num_types", C.c_int32), ("num_type_mlps", C.c_int32),
                ("skip", C.c_int32), ("steps", C.c_int32), ("aux_loss_steps", C.c_int32), ("aggr", C.c_int32),
                ("attn", C.c_int32), ("has_update_mlp", C.c_int32), ("update_hier", C.c_int32), ("num_classes", C.c_int32),
                ("precision", C.c_int32),
                ("node_emb", Mlp), ("edge_emb", Mlp), ("edge_head", Mlp), ("node_head", Mlp), ("class_head", Mlp),
                ("w1_dst", C.c_void_p), ("w1_src", C.c_void_p), ("w1_e0", C.c_void_p), ("w1_e", C.c_void_p),
                ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p), ("wm_x", C.c_void_p),
                ("wm_e", C.c_void_p), ("bm", C.c_void_p), ("wa", C.c_void_p), ("ba", C.c_void_p),
                ("wu", C.c_void_p), ("hier", C.c_void_p), ("bu", C.c_void_p),
                ("tc_w1_e", C.c_void_p), ("tc_w2", C.c_void_p), ("tc_wm_e", C.c_void_p), ("tc_wtab", C.c_void_p),
                ("tc_wu", C.c_void_p), ("tc_wnemb", C.c_void_p), ("tc_wemb", C.c_void_p), ("tc_w1_e0", C.c_void_p), ("tc_wheads", C.c_void_p), ("tc_wh1", C.c_void_p),
                ("tc_wh2", C.c_void_p),
                ("edge_logits", C.c_void_p), ("node_logits", C.c_void_p), ("class_logits", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_uint64)]


class MlpTrain(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("dims", C.c_int32 * (MAX_LAYERS + 1)), ("relu", C.c_int32 * MAX_LAYERS),
                ("bn", C.c_int32 * MAX_LAYERS), ("w", C.c_int64 * MAX_LAYERS), ("b", C.c_int64 * MAX_LAYERS),
                ("gamma", C.c_int64 * MAX_LAYERS), ("beta", C.c_int64 * MAX_LAYERS),
                ("running_mean", C.c_void_p * MAX_LAYERS), ("running_var", C.c_void_p * MAX_LAYERS)]


class MpnTrainParams(C.Structure):
    _fields_ = [("num_nodes", C.c_int64), ("num_edges", C.c_int64), ("x", C.c_void_p), ("edge_attr", C.c_void_p),
                ("edge_index", C.c_void_p), ("dim", C.c_int32), ("skip", C.c_int32), ("steps", C.c_int32),
                ("aux_loss_steps", C.c_int32), ("aggr", C.c_int32), ("has_update_mlp", C.c_int32),
                ("num_classes", C.c_int32), ("params", C.c_void_p), ("grads", C.c_void_p),
                ("node_emb", MlpTrain), ("edge_emb", MlpTrain), ("edge_head", MlpTrain), ("node_head", MlpTrain),
                ("class_head", MlpTrain),
                ("w1", C.c_int64), ("b1", C.c_int64), ("w2", C.c_int64), ("b2", C.c_int64), ("wm", C.c_int64),
                ("bm", C.c_int64), ("wu", C.c_int64), ("bu", C.c_int64),
                ("edge_logits", C.c_void_p), ("node_logits", C.c_void_p), ("class_logits", C.c_void_p),
                ("d_edge_logits", C.c_void_p), ("d_node_logits", C.c_void_p), ("d_class_logits", C.c_void_p),
                ("grad_x", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_uint64),
                ("per_type", C.c_int32), ("num_types", C.c_int32), ("attn", C.c_int32), ("reserved_", C.c_int32),
                ("node_types", C.c_void_p), ("wm_type_stride", C.c_int64), ("wa", C.c_int64), ("ba", C.c_int64)]


class MatchParams(C.Structure):
    _fields_ = [("batch", C.c_int32), ("method", C.c_int32), ("use_neighbours", C.c_int32), ("num_threads", C.c_int32),
                ("matching_radius", C.c_float), ("inclusion_radius", C.c_float), ("sim", C.c_void_p),
                ("sim_stride_b", C.c_int64), ("sim_stride_g", C.c_int64), ("max_gt", C.c_int32), ("max_det", C.c_int32),
                ("num_gt", C.c_void_p), ("num_det", C.c_void_p), ("gt_type", C.c_void_p), ("det_type", C.c_void_p),
                ("cap", C.c_int32), ("match_row", C.c_void_p), ("match_col", C.c_void_p), ("num_match", C.c_void_p),
                ("ambiguous", C.c_void_p), ("node_offsets", C.c_void_p), ("gt_person", C.c_void_p),
                ("node_person", C.c_void_p), ("node_class", C.c_void_p), ("node_label", C.c_void_p),
                ("node_ambiguous", C.c_void_p)]


class LabelArgsParams(C.Structure):
    _fields_ = [("batch", C.c_int32), ("max_persons", C.c_int32), ("num_joints", C.c_int32), ("num_threads", C.c_int32),
                ("clamp_max", C.c_float), ("min_arg", C.c_float), ("det", C.c_void_p), ("node_offsets", C.c_void_p),
                ("gt", C.c_void_p), ("factors", C.c_void_p), ("max_gt", C.c_int32), ("max_det", C.c_int32),
                ("arg", C.c_void_p), ("num_gt", C.c_void_p), ("gt_type", C.c_void_p), ("gt_person", C.c_void_p),
                ("det_type", C.c_void_p)]


CC_METHODS = {"GAEC": 0, "threshold": 1, "greedy": 2}


class GroupParams(C.Structure):
    _fields_ = [("batch", C.c_int32), ("num_joints", C.c_int32), ("num_nodes", C.c_int64), ("num_edges", C.c_int64),
                ("node_threshold", C.c_float), ("cc_method", C.c_int32), ("edge_threshold", C.c_float),
                ("node_offsets", C.c_void_p), ("edge_offsets", C.c_void_p),
                ("edge_index", C.c_void_p), ("joint_det", C.c_void_p), ("node_logits", C.c_void_p),
                ("edge_logits", C.c_void_p), ("class_logits", C.c_void_p), ("person_labels", C.c_void_p),
                ("num_components", C.c_void_p), ("num_kept_edges", C.c_void_p), ("max_persons", C.c_int32),
                ("max_nodes_per_image", C.c_int32), ("persons", C.c_void_p),
                ("num_persons", C.c_void_p), ("mutants", C.c_void_p), ("workspace", C.c_void_p),
                ("workspace_bytes", C.c_uint64)]


class RefineParams(C.Structure):
    _fields_ = [("batch", C.c_int32), ("num_joints", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
                ("tag_dim", C.c_int32), ("max_persons", C.c_int32), ("do_refine", C.c_int32), ("do_adjust", C.c_int32),
                ("scoremaps", C.c_void_p), ("tags", C.c_void_p), ("persons", C.c_void_p), ("num_persons", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_uint64)]


# every symbol include/pgmp.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "pgmp_version": (C.c_int, []),
    "pgmp_last_error": (C.c_char_p, []),
    "pgmp_kernel_launches": (C.c_uint64, []),
    "pgmp_profile_enable": (None, [C.c_int]),
    "pgmp_profile_collect": (C.c_int, [C.c_char_p, C.c_int]),
    "pgmp_gc_workspace_bytes": (C.c_uint64, [C.POINTER(GcParams)]),
    "pgmp_gc_detect": (C.c_int, [C.POINTER(GcParams), C.c_void_p, C.c_void_p]),
    "pgmp_gc_detect_fused": (C.c_int, [C.POINTER(GcParams), C.POINTER(GcAssembly), C.c_void_p, C.c_void_p]),
    "pgmp_gc_gather_stage_tags": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                            C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "pgmp_gc_emit": (C.c_int, [C.POINTER(GcParams), C.POINTER(GcOutputs), C.c_void_p]),
    "pgmp_gc_gather_conv": (C.c_int, [C.POINTER(GatherConvParams), C.c_void_p]),
    "pgmp_gc_gather_conv_patches": (C.c_int, [C.POINTER(GatherConvParams), C.c_void_p, C.c_void_p]),
    "pgmp_gc_gather_conv_backward": (C.c_int, [C.POINTER(GatherConvParams), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "pgmp_gc_assemble_scoremaps": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                             C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pgmp_gc_gather_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64,
                                          C.c_int64, C.c_int64, C.c_int64, C.c_void_p]),
    "pgmp_selftest_umma": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pgmp_selftest_umma_ts": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pgmp_mpn_workspace_bytes": (C.c_uint64, [C.POINTER(MpnParams)]),
    "pgmp_mpn_forward": (C.c_int, [C.POINTER(MpnParams), C.c_void_p]),
    "pgmp_mpn_train_workspace_bytes": (C.c_uint64, [C.POINTER(MpnTrainParams)]),
    "pgmp_mpn_train_forward": (C.c_int, [C.POINTER(MpnTrainParams), C.c_void_p]),
    "pgmp_mpn_train_backward": (C.c_int, [C.POINTER(MpnTrainParams), C.c_void_p]),
    "pgmp_group_workspace_bytes": (C.c_uint64, [C.POINTER(GroupParams)]),
    "pgmp_group_persons": (C.c_int, [C.POINTER(GroupParams), C.c_void_p]),
    "pgmp_refine_workspace_bytes": (C.c_uint64, [C.POINTER(RefineParams)]),
    "pgmp_refine_persons": (C.c_int, [C.POINTER(RefineParams), C.c_void_p]),
    "pgmp_match_labels": (C.c_int, [C.POINTER(MatchParams)]),
    "pgmp_label_similarity_args": (C.c_int, [C.POINTER(LabelArgsParams)]),
    "pgmp_linear_sum_assignment": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
}

_lib = None
_lock = threading.Lock()


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def source_hash():
    """sha256 over the build flags and every file the library is compiled from."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for d in sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + sorted(glob.glob(os.path.join(INCLUDE, "*.h"))):
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


HASH_PATH = LIB_PATH + ".srchash"


def _stale():
    """The library is stale when the hash it was built from differs from the sources' (content, not mtimes:
    a snapshot copied to another box keeps contents, not timestamps)."""
    if not os.path.exists(LIB_PATH) or not os.path.exists(HASH_PATH):
        return True
    with open(HASH_PATH) as f:
        return f.read().strip() != source_hash()


def build(force=False, verbose=False):
    """Compile csrc/*.cu for sm_100a into libpgmp.so (cross-compiles without a GPU)."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libpgmp.so")
    objs = []
    build_dir = os.path.join(PKG_DIR, "build")
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(build_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if True:      # a stale library is rebuilt from every source (12 s): object files are not trusted across boxes
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), out))
    cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed: %s\n%s" % (" ".join(cmd), r.stdout))
    with open(HASH_PATH, "w") as f:
        f.write(source_hash() + "\n")
    return LIB_PATH


def lib():
    """The loaded library; raises if it is missing and cannot be built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if _stale() and (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
                build()
            if not os.path.exists(LIB_PATH):
                raise RuntimeError("libpgmp.so is missing; build it with `python -c 'import __graft_entry__ as g; "
                                   "g.build()'` (needs nvcc). pgmp_b200 has no CPU fallback.")
            handle = C.CDLL(LIB_PATH)
            for name, (res, args) in SYMBOLS.items():
                fn = getattr(handle, name)
                fn.restype, fn.argtypes = res, args
            if handle.pgmp_version() != 100:
                raise RuntimeError("libpgmp.so version mismatch: rebuild")
            _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError("libpgmp: %s (code %d)" % (lib().pgmp_last_error().decode(), rc))


def kernel_launches():
    return int(lib().pgmp_kernel_launches())


def profile(on):
    lib().pgmp_profile_enable(1 if on else 0)


def profile_collect():
    """{kernel name: (launches, total_ms)} since profiling was enabled / last collected."""
    buf = C.create_string_buffer(1 << 16)
    lib().pgmp_profile_collect(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.rsplit(" ", 2)
        out[name.strip("()")] = (int(cnt), float(ms))
    return out


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def current_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(t, name, dtype=None):
    import torch
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: pgmp_b200 has no CPU path" % name)
    if dtype is not None and t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    return t
