"""Case tables shared by ``make_golden.py`` (writes fixtures from the reference) and the tests."""

# name -> (synthetic input kwargs, GC config overrides)
GC_CASES = {
    # exactly K per joint (bench-style config), kNN-50 and fully connected
    "knn_small": (dict(batch=2, num_joints=17, size=128, k=10, persons=3), dict(k=10, graph_type="knn")),
    "fully_small": (dict(batch=2, num_joints=17, size=128, k=10, persons=3), dict(k=10, graph_type="fully")),
    # threshold extras appended after the top-k block; 3x3 pooling
    "thr_extras": (dict(batch=2, num_joints=17, size=128, k=10, persons=3),
                   dict(k=5, graph_type="knn", DETECT_THRESHOLD=0.5, POOL_KERNEL_SIZE=3)),
    # no-threshold path: top-20, scores + 1e-10
    "no_threshold": (dict(batch=1, num_joints=17, size=128, k=12, persons=3),
                     dict(k=5, graph_type="knn", DETECT_THRESHOLD=2.0)),
    # crowd mask multiplies the NMS map
    "mask_crowds": (dict(batch=2, num_joints=17, size=128, k=10, persons=3),
                    dict(k=10, graph_type="knn", MASK_CROWDS=True)),
    # non-square map, norm = 160 (inexact fp32 division), un-normalised variant too
    "rect_norm": (dict(batch=1, num_joints=17, size=96, width=160, k=8, persons=2), dict(k=8, graph_type="knn")),
    "rect_nonorm": (dict(batch=1, num_joints=17, size=96, width=160, k=8, persons=2),
                    dict(k=8, graph_type="fully", NORM_NODE_DISTANCE=False)),
    # N <= 51 -> kNN graph is complete
    "tiny_complete": (dict(batch=2, num_joints=4, size=64, k=5, persons=2), dict(k=5, graph_type="knn")),
    # CrowdPose-shaped: 14 joints, 60 candidates per joint
    "crowdpose": (dict(batch=1, num_joints=14, size=256, k=60, persons=20), dict(k=60, graph_type="knn")),
    # other edge-feature sets
    "feat_position": (dict(batch=1, num_joints=17, size=128, k=10, persons=3),
                      dict(k=10, graph_type="knn", EDGE_FEATURES_TO_USE=["position"])),
    "feat_type": (dict(batch=1, num_joints=17, size=128, k=10, persons=3),
                  dict(k=10, graph_type="knn", EDGE_FEATURES_TO_USE=["connection_type"])),
    # BASELINE.json configs[0]: single 512x512 image, 30 per joint, kNN-50 (large arrays stored as digests)
    "config1_512": (dict(batch=1, num_joints=17, size=512, k=30, persons=8), dict(k=30, graph_type="knn")),
}

# name -> (GC case providing the graph, MPN config maker name, overrides, weight seed)
MPN_CASES = {
    "flagship": ("knn_small", "flagship_mpn_config", dict(), 1),
    "flagship_aux2": ("knn_small", "flagship_mpn_config", dict(AUX_LOSS_STEPS=2, STEPS=4), 2),
    "agnostic_max": ("knn_small", "agnostic_mpn_config", dict(), 3),
    "agnostic_add_noskip_upd": ("knn_small", "agnostic_mpn_config",
                                dict(AGGR="add", SKIP=False, USE_NODE_UPDATE_MLP=True, STEPS=3), 4),
    "agnostic_mean": ("tiny_complete", "agnostic_mpn_config", dict(AGGR="mean", STEPS=3, NUM_JOINTS=4,
                                                                   EDGE_INPUT_DIM=6), 5),
    "pertype_vanilla_add": ("knn_small", "flagship_mpn_config", dict(AGGR_SUB="None", AGGR="add", STEPS=3), 6),
    "pertype_vanilla_max": ("knn_small", "flagship_mpn_config", dict(AGGR_SUB="None", AGGR="max", STEPS=3), 7),
    "pertype_attn_per_type": ("knn_small", "flagship_mpn_config", dict(AGGR_SUB="node_edge_attn_per_type", STEPS=3), 8),
    "flagship_fully": ("fully_small", "flagship_mpn_config", dict(STEPS=2), 9),
    "flagship_crowdpose": ("crowdpose", "flagship_mpn_config", dict(STEPS=2, NUM_JOINTS=14, EDGE_INPUT_DIM=16), 10),
    "flagship_config1": ("config1_512", "flagship_mpn_config", dict(), 11),
    "pertype_hierarch_mlp": ("knn_small", "flagship_mpn_config", dict(UPDATE_TYPE="hierarch_mlp", STEPS=3), 12),
}

# arrays at or above this many bytes are stored as sha256 digests instead of values
DIGEST_BYTES = 300_000


def gc_config_for(pgmp_config, overrides):
    o = dict(overrides)
    return pgmp_config.bench_gc_config(k=o.pop("k"), graph_type=o.pop("graph_type"), **o)


def mpn_config_for(pgmp_config, maker, overrides):
    o = dict(overrides)
    nj = o.get("NUM_JOINTS", 17)
    cfg = getattr(pgmp_config, maker)(nj, **o)
    if nj != 17:                                    # CLASS head width follows the dataset's joint count
        cfg.CLASS.OUTPUT_SIZES = [64, 32, nj]
    return cfg


# ---- training-mode cases (round-2 groundwork): name -> (GC case, MPN config maker, overrides, weight seed)
TRAIN_CASES = {
    # class_agnostic_end2end/model_57_1_0.yaml shape: agnostic MPLayer, max aggregation, skip
    "agnostic_max": ("knn_small", "agnostic_mpn_config", dict(STEPS=3, AUX_LOSS_STEPS=1), 41),
    "flagship": ("knn_small", "flagship_mpn_config", dict(STEPS=2), 42),
}


def train_loss_weights(shapes, seed):
    """Fixed random coefficients of the scalar test loss L = sum_k <c_k, pred_k>."""
    import numpy as np
    rng = np.random.default_rng(1000 + seed)
    return [rng.standard_normal(s).astype(np.float32) for s in shapes]


def sample_indices(n, name):
    """64 fixed positions of a flattened parameter gradient (all of it when smaller)."""
    import hashlib
    import numpy as np
    if n <= 64:
        return np.arange(n)
    rng = np.random.default_rng(int.from_bytes(hashlib.sha256(name.encode()).digest()[:4], "little"))
    return np.sort(rng.choice(n, 64, replace=False))


# ---- pose-assembly tail (refine / adjust): name -> (batch, joints, height, width, tag dim or None, persons per image, seed)
REFINE_CASES = {
    "coco_t1": (2, 17, 64, 80, None, (5, 3), 1),
    "tags2": (1, 17, 48, 48, 2, (4,), 2),
    "crowd14": (1, 14, 40, 56, None, (9,), 3),
}


def refine_inputs(name):
    """Seeded inputs of a REFINE case: scoremaps [B,J,H,W] f32, tags [B,J,H,W(,T)] f32, keypoints: list of [P,J,3] f64
    (x, y, score) with about a third of the joints missing (score 0, position = the mean of the detected ones as
    ``fill_mean`` leaves them, Utils.py:1469-1471)."""
    import numpy as np
    B, J, H, W, T, persons, seed = REFINE_CASES[name]
    rng = np.random.default_rng(7000 + seed)
    sm = rng.uniform(0, 0.05, (B, J, H, W)).astype(np.float32)
    for b in range(B):
        for j in range(J):
            for _ in range(6):
                y, x = int(rng.integers(1, H - 1)), int(rng.integers(1, W - 1))
                sm[b, j, y - 1:y + 2, x - 1:x + 2] += rng.uniform(0.1, 0.9, (3, 3)).astype(np.float32)
    shape = (B, J, H, W) if T is None else (B, J, H, W, T)
    tags = (rng.standard_normal(shape) * 2.0).astype(np.float32)
    kps = []
    for b in range(B):
        k = np.zeros((persons[b], J, 3), np.float64)
        for p in range(persons[b]):
            det = rng.uniform(size=J) > 0.35
            det[int(rng.integers(0, J))] = True
            k[p, :, 0] = rng.integers(0, W, J) + rng.choice([0.0, 0.25, 0.5], J)
            k[p, :, 1] = rng.integers(0, H, J) + rng.choice([0.0, 0.25, 0.5], J)
            k[p, :, 2] = np.where(det, rng.uniform(0.1, 1.0, J), 0.0)
            k[p, ~det, :2] = k[p, det, :2].mean(axis=0)
        kps.append(k)
    return sm, tags, kps


# ---- BASELINE.json configs[2..3] at full per-image size, 10 steps:
#      name -> (synthetic input kwargs, GC config overrides, flagship MPN overrides, weight seed)
FULL_CASES = {
    "w48_640_fully": (dict(batch=1, num_joints=17, size=640, k=30, persons=8), dict(k=30, graph_type="fully"), dict(), 21),
    "crowdpose_knn": (dict(batch=1, num_joints=14, size=512, k=60, persons=20), dict(k=60, graph_type="knn"),
                      dict(NUM_JOINTS=14, EDGE_INPUT_DIM=16), 22),
    "crowdpose_fully": (dict(batch=1, num_joints=14, size=512, k=60, persons=20), dict(k=60, graph_type="fully"),
                        dict(NUM_JOINTS=14, EDGE_INPUT_DIM=16), 23),
}
FULL_EDGE_SAMPLE = 8192


def full_edge_sample(n, name):
    """The fixed edge positions whose logits a full-size fixture stores."""
    import hashlib
    import numpy as np
    if n <= FULL_EDGE_SAMPLE:
        return np.arange(n)
    rng = np.random.default_rng(int.from_bytes(hashlib.sha256(("full_" + name).encode()).digest()[:4], "little"))
    return np.sort(rng.choice(n, FULL_EDGE_SAMPLE, replace=False))


# ---- training branch of the graph constructor (labels): name -> (synthetic input kwargs, GC config overrides)
LABEL_CASES = {
    "m6": (dict(batch=2, num_joints=17, size=128, k=10, persons=3),
           dict(k=10, graph_type="knn", EDGE_LABEL_METHOD=6, MATCHING_RADIUS=0.5, INCLUSION_RADIUS=0.75)),
    "m6_neigh_bg": (dict(batch=2, num_joints=17, size=128, k=10, persons=3),
                    dict(k=10, graph_type="knn", EDGE_LABEL_METHOD=6, MATCHING_RADIUS=0.3, INCLUSION_RADIUS=0.35,
                         USE_NEIGHBOURS=True, WITH_BACKGROUND=True, POOL_KERNEL_SIZE=3)),
    "m4_neigh": (dict(batch=2, num_joints=17, size=128, k=10, persons=3),
                 dict(k=10, graph_type="fully", EDGE_LABEL_METHOD=4, MATCHING_RADIUS=0.1, INCLUSION_RADIUS=0.3,
                      USE_NEIGHBOURS=True, POOL_KERNEL_SIZE=3)),
    "m6_thr": (dict(batch=2, num_joints=17, size=128, k=10, persons=3),
               dict(k=5, graph_type="knn", EDGE_LABEL_METHOD=6, MATCHING_RADIUS=0.5, DETECT_THRESHOLD=0.3)),
}
LABEL_SLOTS = dict(edge_labels=3, node_labels=4, node_classes=5, label_mask=8, label_mask_node=9, class_mask=10, node_persons=13)


def label_inputs(name):
    """Seeded inputs of a LABEL case: the synthetic batch, ``joints_gt [B, 30, J, 3]`` and ``factors [B, 30, J]``."""
    import numpy as np
    import pgmp_b200.synthetic as synthetic
    inp_kw, _ = LABEL_CASES[name]
    data = synthetic.synth_batch(**inp_kw)
    gts, facs = zip(*[synthetic.synth_joints_gt(b, inp_kw["num_joints"], inp_kw["size"], inp_kw["k"], inp_kw["persons"])
                      for b in range(inp_kw["batch"])])
    return data, np.stack(gts), np.stack(facs)
