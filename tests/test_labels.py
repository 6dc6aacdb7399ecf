"""Training branch of the graph constructor: the label slots of the 15-tuple against the UNMODIFIED reference
(``tests/golden/labels_*.npz``, written by ``make_golden_labels.py``).  The matching runs on the host with the
reference's own operations, the per-edge labels are integer compares: everything bit-exact."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import oracle
import pgmp_b200
from cases import LABEL_CASES, LABEL_SLOTS, gc_config_for, label_inputs
from helpers import golden
from pgmp_b200.graph_constructor import labels as L


def _check(gold, got):
    for key in LABEL_SLOTS:
        if bool(gold[key + "__none"]):
            assert got[key] is None, key
            continue
        a = got[key].cpu().numpy()
        assert a.dtype == gold[key].dtype and a.shape == gold[key].shape, (key, a.dtype, gold[key].dtype, a.shape)
        assert np.array_equal(a, gold[key]), f"{key}: {np.sum(a != gold[key])} mismatches"


@pytest.mark.parametrize("name", list(LABEL_CASES))
def test_label_construction_on_the_oracle_graph(name):
    """``labels.build_labels`` on the oracle's graph (CPU tensors): matching + per-edge labels against the reference."""
    inp_kw, cfg_over = LABEL_CASES[name]
    data, gt, factors = label_inputs(name)
    cfg = gc_config_for(pgmp_b200.config, cfg_over)
    g = oracle.gc.construct_graph(data["scoremaps"], data["tagmaps"], data["features"], cfg, inp_kw["num_joints"])
    gold = golden("labels_" + name)
    assert np.array_equal(g["joint_det"], gold["joint_det"]) and g["edge_index"].shape[1] == int(gold["num_edges"])
    gc = SimpleNamespace(edge_label_method=cfg.EDGE_LABEL_METHOD, joints_gt=torch.from_numpy(gt),
                         factor_list=torch.from_numpy(factors), scoremaps=torch.from_numpy(data["scoremaps"]),
                         matching_radius=cfg.MATCHING_RADIUS, inclusion_radius=cfg.INCLUSION_RADIUS,
                         include_neighbouring_keypoints=cfg.USE_NEIGHBOURS, with_background_class=cfg.WITH_BACKGROUND,
                         num_joints=inp_kw["num_joints"])
    lab = L.build_labels(gc, torch.from_numpy(g["joint_det"]), torch.from_numpy(g["edge_index"]),
                         torch.from_numpy(g["batch_index"]), g["num_nodes"].tolist())
    _check(gold, lab)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(LABEL_CASES))
def test_construct_graph_with_ground_truth_matches_reference(name):
    """The drop-in call the training loop makes (``PoseEstimation.forward``: ``joints_gt=keypoints, factor_list=factors``)."""
    from pgmp_b200.graph_constructor import get_graph_constructor
    inp_kw, cfg_over = LABEL_CASES[name]
    data, gt, factors = label_inputs(name)
    cfg = gc_config_for(pgmp_b200.config, cfg_over)
    dev = "cuda:0"
    t = {k: torch.from_numpy(v).to(dev) for k, v in data.items()}
    ret = get_graph_constructor(cfg, scoremaps=t["scoremaps"], tagmaps=t["tagmaps"], features=t["features"],
                                joints_gt=torch.from_numpy(gt).to(dev), factor_list=torch.from_numpy(factors).to(dev),
                                masks=None, device=dev, testing=True, heatmaps=None,
                                num_joints=inp_kw["num_joints"]).construct_graph()
    gold = golden("labels_" + name)
    assert np.array_equal(ret[7].cpu().numpy(), gold["joint_det"])
    _check(gold, {k: ret[slot] for k, slot in LABEL_SLOTS.items()})


@pytest.mark.gpu
def test_node_dropout_removes_only_positive_nodes_and_renumbers():
    from pgmp_b200.graph_constructor import get_graph_constructor
    name = "m6"
    inp_kw, cfg_over = LABEL_CASES[name]
    data, gt, factors = label_inputs(name)
    dev = "cuda:0"
    t = {k: torch.from_numpy(v).to(dev) for k, v in data.items()}
    kw = dict(scoremaps=t["scoremaps"], tagmaps=t["tagmaps"], features=t["features"], joints_gt=torch.from_numpy(gt).to(dev),
              factor_list=torch.from_numpy(factors).to(dev), masks=None, device=dev, heatmaps=None, num_joints=inp_kw["num_joints"])
    full = get_graph_constructor(gc_config_for(pgmp_b200.config, cfg_over), testing=False, **kw).construct_graph()
    torch.manual_seed(0)
    gc = get_graph_constructor(gc_config_for(pgmp_b200.config, dict(cfg_over, NODE_DROPOUT=0.5)), testing=False, **kw)
    ret = gc.construct_graph()
    n_pos, n_neg = int(full[4].sum()), int((full[4] == 0).sum())
    assert int((ret[4] == 0).sum()) == n_neg and 0 < int(ret[4].sum()) < n_pos            # negatives all stay
    N, E = ret[0].shape[0], ret[2].shape[1]
    assert ret[7].shape[0] == N == ret[11].shape[0] == ret[12].shape[0] == ret[13].shape[0] and ret[1].shape[0] == E == ret[3].shape[0]
    assert int(ret[2].max()) < N and int(gc.num_nodes_per_image.sum()) == N and int(gc.num_edges_per_image.sum()) == E
    # the surviving edges carry the labels they had: positive iff both ends belong to the same person
    ps, pd = ret[13][ret[2][0]], ret[13][ret[2][1]]
    assert torch.equal(ret[3], ((ps == pd) & (ps >= 0)).float())
    assert bool((ret[12][ret[2][0]] == ret[12][ret[2][1]]).all())


def test_native_assignment_equals_scipy_on_tie_heavy_matrices():
    """csrc/match.cu restates scipy.optimize.linear_sum_assignment (the thresholded similarity matrices are mostly zeros,
    so the tie rule decides): identical rows / columns on random dense, sparse, integer and tall matrices."""
    from scipy.optimize import linear_sum_assignment
    import pgmp_b200._native as nv
    lib = nv.lib()
    rng = np.random.default_rng(5)
    for trial in range(600):
        nr, nc = int(rng.integers(1, 40)), int(rng.integers(1, 60))
        kind = trial % 4
        if kind == 0:
            cost = rng.random((nr, nc))
        elif kind == 1:
            cost = rng.random((nr, nc))
            cost[cost < 0.8] = 0
        elif kind == 2:
            cost = rng.integers(0, 3, (nr, nc)).astype(np.float64)
        else:
            cost = rng.random((nr, nc)).astype(np.float32).astype(np.float64)
            cost[rng.random((nr, nc)) < 0.95] = 0
        cost = np.ascontiguousarray(cost)
        maximize = bool(trial % 2)
        r, c = linear_sum_assignment(cost, maximize=maximize)
        rr, cc = np.zeros(min(nr, nc), np.int64), np.zeros(min(nr, nc), np.int64)
        nv.check(lib.pgmp_linear_sum_assignment(cost.ctypes.data, nr, nc, int(maximize), rr.ctypes.data, cc.ctypes.data))
        assert np.array_equal(r, rr) and np.array_equal(c, cc), (trial, nr, nc)


@pytest.mark.parametrize("method,neigh", [(4, False), (4, True), (6, False), (6, True)])
def test_native_matching_equals_the_per_image_restatement(method, neigh):
    """``pgmp_match_labels`` (whole batch, native) against ``oracle/labels.py`` (per image, torch + scipy: the reference's
    operations) on crowded random scenes: more annotated joints than candidates of a type, ties, empty images."""
    import oracle.labels as OL
    rng = np.random.default_rng(100 + method + int(neigh))
    B, J, S = 6, 5, 64
    dets, counts, gts, facs = [], [], [], []
    for b in range(B):
        n = 0 if b == 4 else int(rng.integers(3, 40))
        d = np.stack([rng.integers(0, S, n), rng.integers(0, S, n), rng.integers(0, J, n)], 1).astype(np.int64).reshape(n, 3)
        gt = np.zeros((7, J, 3), np.float32)
        gt[:, :, :2] = rng.random((7, J, 2)) * S
        gt[:, :, 2] = rng.random((7, J)) < (0.0 if b == 2 else 0.7)
        if n and b != 2:          # some annotated joints sit exactly on candidates, some candidates are shared
            for k in range(min(n, 6)):
                gt[k % 7, d[k, 2], :2] = d[k, :2] + rng.integers(-1, 2, 2)
        dets.append(d)
        counts.append(n)
        gts.append(gt)
        facs.append((rng.random((7, J)) * 40 + 4).astype(np.float32))
    det_all = torch.from_numpy(np.concatenate(dets, 0))
    gc = SimpleNamespace(edge_label_method=method, joints_gt=torch.from_numpy(np.stack(gts)), factor_list=torch.from_numpy(np.stack(facs)),
                         scoremaps=torch.zeros(B, J, S, S), matching_radius=0.1, inclusion_radius=0.35,
                         include_neighbouring_keypoints=neigh, with_background_class=False, num_joints=J)
    N = det_all.shape[0]
    edge_index = torch.from_numpy(np.stack([rng.integers(0, N, 50), rng.integers(0, N, 50)]))
    batch_index = torch.repeat_interleave(torch.arange(B), torch.tensor(counts))
    lab = L.build_labels(gc, det_all, edge_index, batch_index, counts)
    off = 0
    for b in range(B):
        n = counts[b]
        nodes, persons, joints, amb = OL.match_image(det_all[off:off + n], gc.joints_gt[b], gc.factor_list[b], method, S, 0.1, 0.35, neigh)
        want_person = np.full(n, -1, np.int64)
        want_person[nodes.numpy()] = persons.numpy()
        assert np.array_equal(lab["node_persons"][off:off + n].numpy(), want_person), b
        want_label = np.zeros(n, np.float32)
        want_label[nodes.numpy()] = 1
        assert np.array_equal(lab["node_labels"][off:off + n].numpy(), want_label), b
        if method == 6:
            want_class = np.zeros(n, np.int64)
            want_class[nodes.numpy()] = joints.numpy()
            assert np.array_equal(lab["node_classes"][off:off + n].numpy(), want_class), b
        if neigh:
            assert np.array_equal(lab["label_mask_node"][off:off + n].numpy(), 1.0 - amb.astype(np.float32)), b
        off += n
