"""Shared test helpers: golden fixtures (written by tests/golden/make_golden.py from the
unmodified reference), synthetic inputs, comparison utilities."""
import functools
import hashlib
import os

import numpy as np

import pgmp_b200
import pgmp_b200.synthetic as synthetic
from cases import GC_CASES, MPN_CASES, gc_config_for, mpn_config_for  # noqa: F401

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GC_KEYS = ["x", "edge_attr", "edge_index", "joint_det", "joint_scores", "batch_index", "joint_tags"]

# north_star tolerance: logits within 1e-3 relative error.  "Relative" is measured against the
# tensor's scale: max|a - b| <= tol * max|b| and ||a - b||_2 <= tol * ||b||_2.
FP32_TOL = 2e-5      # fp32 mode: only the summation order differs from the reference
FP32_TOL_FULL = 5e-5 # fp32 mode on the full-size configs (sums over up to 839 edges per target, 10 steps)
LOGIT_TOL = 1e-3     # tensor-core mode (north_star)


def golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def assert_matches_golden(gold, key, arr, exact=True, tol=None):
    """Compare ``arr`` with fixture entry ``key`` (stored as values or as a sha256 digest)."""
    arr = np.ascontiguousarray(arr)
    if key in gold:
        ref = gold[key]
        assert arr.shape == ref.shape, (key, arr.shape, ref.shape)
        assert arr.dtype == ref.dtype, (key, arr.dtype, ref.dtype)
        if exact:
            assert np.array_equal(arr, ref), f"{key}: {np.sum(arr != ref)} mismatches"
        else:
            assert_close(arr, ref, tol, key)
    else:
        assert tuple(gold[key + "__shape"]) == arr.shape, (key, arr.shape)
        assert exact, "digest entries are bit-exact quantities"
        assert np.array_equal(sha(arr), gold[key + "__sha256"]), f"{key}: digest mismatch"


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-30) if b.size else 1.0
    linf = np.abs(a - b).max() / scale if b.size else 0.0
    l2 = np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30) if b.size else 0.0
    return linf, l2


def assert_close(a, b, tol, what=""):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    linf, l2 = rel_err(a, b)
    assert linf <= tol and l2 <= tol, f"{what}: rel linf {linf:.3e} l2 {l2:.3e} > {tol:.1e}"


@functools.lru_cache(maxsize=4)
def gc_inputs(name):
    inp_kw, cfg_over = GC_CASES[name]
    data = synthetic.synth_batch(**inp_kw)
    cfg = gc_config_for(pgmp_b200.config, cfg_over)
    return data, cfg, inp_kw["num_joints"]
