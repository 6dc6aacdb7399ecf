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


def test_filter_and_fill_follows_pred_to_ann():
    """Host steps of pred_to_ann (Utils.py:1463-1471), stated directly."""
    from pgmp_b200.Utils import filter_and_fill
    rng = np.random.default_rng(2)
    p = np.zeros((4, 17, 3))
    p[:, :, :2] = rng.uniform(0, 100, (4, 17, 2))
    p[:, :, 2] = np.where(rng.uniform(size=(4, 17)) > 0.4, rng.uniform(0.05, 1.0, (4, 17)), 0.0)
    p[2, :, 2] = np.where(p[2, :, 2] > 0, 0.2, 0.0)                  # best score 0.2: filtered
    p[:, 3, 2] = np.maximum(p[:, 3, 2], 0.1)
    want = p.copy()
    keep = want[:, :, 2].max(axis=1) > 0.25
    want = want[keep]
    for i in range(len(want)):
        want[i, want[i, :, 2] == 0, :2] = want[i, want[i, :, 2] != 0, :2].mean(axis=0)
    got = filter_and_fill(p, with_filter=True, fill_mean=True)
    assert got.shape[0] == 3 and np.array_equal(got, want)
    assert np.array_equal(filter_and_fill(p, with_filter=False, fill_mean=False), p)
    assert filter_and_fill(np.array([]), True, True) is None
    low = p.copy()
    low[:, :, 2] *= 0.1
    assert filter_and_fill(low, with_filter=True) is None


@pytest.mark.gpu
def test_persons_from_groups_chain():
    """group_persons output -> filter / fill (host) -> refine + adjust (device) against the oracle chain."""
    from pgmp_b200.Utils import filter_and_fill, persons_from_groups
    sm, tags, kps = refine_inputs("coco_t1")
    raw = []
    for k in kps:                                                    # undo the fixture's fill: arbitrary positions at missing joints
        r = k.copy()
        r[r[:, :, 2] == 0, :2] = 0.0
        raw.append(r)
    groups = [(raw[0], False, None), None]
    got = persons_from_groups(torch.from_numpy(sm).cuda(), torch.from_numpy(tags).cuda(), groups, with_filter=True)
    assert got[1] is None
    want = filter_and_fill(raw[0], True, True)
    want = R.adjust(R.refine(sm[0], tags[0], want), sm[0])
    assert np.array_equal(got[0], want)
