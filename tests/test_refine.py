"""Pose-assembly tail after the grouping: ``refine`` / ``adjust`` (src/Utils/Utils.py:1026-1104,
917-936).  CPU: the numpy oracle against the reference's own functions (tests/golden/refine_*.npz, written by
tests/golden/make_golden_refine.py from the reference source).  GPU: the batched CUDA kernels against the oracle and the
fixtures, bit-exact (float64 coordinates, the arg-max decided by numpy's float32 arithmetic)."""
import os

import numpy as np
import pytest
import torch

import oracle.refine as R
from cases import REFINE_CASES, refine_inputs

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("name", list(REFINE_CASES))
def test_oracle_matches_reference_functions(name):
    sm, tags, kps = refine_inputs(name)
    gold = np.load(os.path.join(HERE, "golden", f"refine_{name}.npz"))
    added = 0
    for b in range(sm.shape[0]):
        r = R.refine(sm[b], tags[b], kps[b])
        assert np.array_equal(r, gold[f"refined_{b}"])
        assert np.array_equal(R.adjust(r, sm[b]), gold[f"adjusted_{b}"])
        assert np.array_equal(R.adjust(kps[b], sm[b]), gold[f"adjusted_only_{b}"])
        added += int((r[:, :, 2] == 0.001).sum())
    assert added > 10


def test_oracle_mean_tag_is_numpy_float32_mean():
    rng = np.random.default_rng(0)
    for n in (1, 7, 8, 13, 17):
        a = (rng.standard_normal((n, 1)) * 3).astype(np.float32)
        assert R.mean_tag(a).dtype == np.float32 and np.array_equal(R.mean_tag(a), np.mean(a, axis=0))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(REFINE_CASES))
def test_cuda_refine_adjust_bit_exact(name):
    from pgmp_b200.Utils import refine_persons
    sm, tags, kps = refine_inputs(name)
    gold = np.load(os.path.join(HERE, "golden", f"refine_{name}.npz"))
    sm_d, tg_d = torch.from_numpy(sm).cuda(), torch.from_numpy(tags).cuda()
    refined = refine_persons(sm_d, tg_d, kps, with_refine=True, adjustment=False)
    both = refine_persons(sm_d, tg_d, kps, with_refine=True, adjustment=True)
    only = refine_persons(sm_d, tg_d, kps, with_refine=False, adjustment=True)
    for b in range(sm.shape[0]):
        assert np.array_equal(refined[b], R.refine(sm[b], tags[b], kps[b])), "refine vs oracle"
        assert np.array_equal(refined[b], gold[f"refined_{b}"]), "refine vs reference"
        assert np.array_equal(both[b], gold[f"adjusted_{b}"]), "refine + adjust vs reference"
        assert np.array_equal(only[b], gold[f"adjusted_only_{b}"]), "adjust vs reference"


@pytest.mark.gpu
def test_cuda_refine_ragged_batch_and_full_size():
    """Images without persons, more persons than one pass holds (8), a 512 x 512 map: against the oracle."""
    from pgmp_b200.Utils import refine_persons
    rng = np.random.default_rng(11)
    B, J, H, W = 3, 17, 512, 512
    sm = rng.uniform(0, 0.05, (B, J, H, W)).astype(np.float32)
    sm[:, :, 100:103, 200:203] += 0.5
    tags = (rng.standard_normal((B, J, H, W)) * 2).astype(np.float32)
    kps = [None, np.zeros((11, J, 3)), np.zeros((0, J, 3))]
    k = kps[1]
    k[:, :, 0], k[:, :, 1] = rng.integers(0, W, (11, J)), rng.integers(0, H, (11, J))
    k[:, :, 2] = np.where(rng.uniform(size=(11, J)) > 0.5, rng.uniform(0.1, 1, (11, J)), 0.0)
    k[:, 0, 2] = 0.7
    out = refine_persons(torch.from_numpy(sm).cuda(), torch.from_numpy(tags).cuda(), kps)
    assert out[0] is None and out[2].shape == (0, J, 3)
    assert np.array_equal(out[1], R.adjust(R.refine(sm[1], tags[1], k), sm[1]))
