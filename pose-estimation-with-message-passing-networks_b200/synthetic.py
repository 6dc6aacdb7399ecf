"""Deterministic synthetic inputs for the grouping path.

Heatmaps follow the reference's ``HeatmapGenerator`` (``src/data/utils.py:30-65``):
sigma = 2 Gaussians combined with ``max``, here with rank-unique amplitudes so
that the per-joint top-k is free of ties, plus U(0, 0.02) noise that keeps
plateaus off zero.  Everything is seeded per image (``1000 + image_index``) and
generated with numpy so the same arrays feed the CPU oracle and the CUDA path.
"""

import numpy as np

SIGMA = 2.0
_R = 6  # render radius: exp(-36 / 8) ~ 1e-2 of the amplitude floor is below the noise


def _gauss_patch():
    ax = np.arange(-_R, _R + 1, dtype=np.float32)
    return np.exp(-(ax[None, :] ** 2 + ax[:, None] ** 2) / np.float32(2 * SIGMA * SIGMA)).astype(np.float32)


def synth_scoremap(image_index, num_joints=17, size=512, k=30, persons=8, noise=0.02, width=None):
    """One image's ``[J, H, W]`` float32 scoremap with >= 2k local maxima per joint."""
    H = size
    W = size if width is None else width
    rng = np.random.default_rng(1000 + image_index)
    out = np.zeros((num_joints, H, W), dtype=np.float32)
    patch = _gauss_patch()
    centres = rng.uniform(0.15, 0.85, size=(persons, 2)) * np.array([W, H])
    n_peaks = 2 * k + 8
    for j in range(num_joints):
        pts = np.empty((n_peaks, 2), dtype=np.int64)
        # the first `persons` peaks: skeleton joints jittered around each person centre
        jit = rng.normal(0.0, 0.1, size=(persons, 2)) * np.array([W, H])
        body = centres + jit
        rnd = rng.uniform(0, 1, size=(n_peaks, 2)) * np.array([W, H])
        rnd[:min(persons, n_peaks)] = body[:min(persons, n_peaks)]
        pts[:, 0] = np.clip(np.rint(rnd[:, 0]), _R, W - 1 - _R)
        pts[:, 1] = np.clip(np.rint(rnd[:, 1]), _R, H - 1 - _R)
        amps = np.linspace(0.2, 0.95, n_peaks, dtype=np.float32)
        rng.shuffle(amps)
        for (px, py), a in zip(pts, amps):
            win = out[j, py - _R:py + _R + 1, px - _R:px + _R + 1]
            np.maximum(win, a * patch, out=win)
    out += rng.uniform(0.0, noise, size=out.shape).astype(np.float32)
    return out


def synth_mask(image_index, height, width):
    """A crowd mask ``[H, W]`` float32 in {0, 1} with one zeroed rectangle (``MASK_CROWDS``)."""
    rng = np.random.default_rng(9000 + image_index)
    m = np.ones((height, width), dtype=np.float32)
    y0, x0 = int(rng.integers(0, height // 2)), int(rng.integers(0, width // 2))
    m[y0:y0 + height // 3, x0:x0 + width // 3] = 0
    return m


def synth_batch(batch, num_joints=17, size=512, k=30, channels=128, persons=8, first_index=0,
                with_features=True, tag_dim=None, width=None):
    """Host-side batch: dict of ``scoremaps [B,J,H,W]``, ``features [B,C,H,W]``,
    ``tagmaps [B,J,H,W]`` (or ``[B,J,H,W,T]``), ``masks [B,H,W]``, all float32."""
    H, W = size, (size if width is None else width)
    sm = np.stack([synth_scoremap(first_index + b, num_joints, size, k, persons, width=width) for b in range(batch)])
    out = {"scoremaps": sm,
           "masks": np.stack([synth_mask(first_index + b, H, W) for b in range(batch)])}
    if with_features:
        feats, tags = [], []
        for b in range(batch):
            rng = np.random.default_rng(5000 + first_index + b)
            feats.append(rng.standard_normal((channels, H, W), dtype=np.float32))
            shape = (num_joints, H, W) if tag_dim is None else (num_joints, H, W, tag_dim)
            tags.append(rng.standard_normal(shape, dtype=np.float32))
        out["features"] = np.stack(feats)
        out["tagmaps"] = np.stack(tags)
    return out


def synth_mpn_state_dict(model, seed=0):
    """Fill an MPN module's parameters and BatchNorm statistics deterministically.

    Values depend only on (seed, tensor name, shape) -- not on module registration
    order -- so the reference model and the drop-in module get identical weights.
    Weights ~ U(+-sqrt(6/fan_in)) (ReLU-preserving scale, so that logits are O(1));
    BatchNorm gets non-trivial affine terms and running statistics.
    """
    import zlib

    import torch

    with torch.no_grad():
        for name, t in sorted(model.state_dict().items()):
            if name.endswith("num_batches_tracked"):
                continue
            g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 31))
            if t.dim() >= 2:
                v = (torch.rand(t.shape, generator=g) * 2 - 1) * (6.0 / t.shape[1]) ** 0.5
            elif name.endswith("running_mean"):
                v = 0.1 * torch.randn(t.shape, generator=g)
            elif name.endswith("running_var"):
                v = 0.5 + torch.rand(t.shape, generator=g)
            elif name.endswith("bias"):
                v = (torch.rand(t.shape, generator=g) * 2 - 1) * 0.1
            else:  # BatchNorm weight
                v = 0.75 + 0.5 * torch.rand(t.shape, generator=g)
            t.copy_(v.to(t.device))
    return model


def synth_group_logits(joint_det, batch_index, edge_index, num_joints=17, persons=6, seed=0):
    """Logits with a planted person partition, for exercising the grouping tail
    (threshold -> multicut -> persons) independently of the MPN weights.

    ~20 % of the nodes are false positives (negative node logit); edges inside a
    planted person get positive logits, all others negative; continuous noise makes
    every multicut weight distinct (no ties in the greedy contraction order).
    """
    rng = np.random.default_rng(7000 + seed)
    n, e = len(joint_det), edge_index.shape[1]
    person = rng.integers(0, persons, size=n)
    person[rng.uniform(size=n) < 0.2] = -1
    node_logits = np.where(person >= 0, 3.0, -4.0) + rng.normal(0, 1.0, size=n)
    src, dst = edge_index
    same = (person[src] == person[dst]) & (person[src] >= 0)
    edge_logits = np.where(same, 2.5, -2.5) + rng.normal(0, 1.0, size=e)
    class_logits = rng.normal(0, 1.0, size=(n, num_joints))
    typ = np.where(rng.uniform(size=n) < 0.9, joint_det[:, 2], rng.integers(0, num_joints, size=n))
    class_logits[np.arange(n), typ] += 4.0
    return dict(node_logits=node_logits.astype(np.float32), edge_logits=edge_logits.astype(np.float32),
                class_logits=class_logits.astype(np.float32), num_joints=num_joints)


def image_subgraph(graph, logits, b):
    """Nodes / edges / logits of image ``b`` with image-local node ids."""
    nsel = graph["batch_index"] == b
    off = int(np.flatnonzero(nsel)[0])
    esel = nsel[graph["edge_index"][0]]
    return dict(joint_det=graph["joint_det"][nsel], edge_index=graph["edge_index"][:, esel] - off,
                node_logits=logits["node_logits"][nsel], edge_logits=logits["edge_logits"][esel],
                class_logits=logits["class_logits"][nsel], num_joints=logits["num_joints"])


def synth_joints_gt(image_index, num_joints=17, size=512, k=30, persons=8, width=None, max_persons=30, drop=0.2):
    """Ground truth that goes with ``synth_scoremap(image_index, ...)``: the planted persons' joints.

    Returns ``joints_gt [max_persons, J, 3]`` float32 (x, y, visibility; the layout ``PoseEstimation.forward`` hands to the
    graph constructor, ``ConstructGraph.py:11``) and ``factors [max_persons, J]`` float32 (the OKS-style ``2 (k_j s)^2``
    denominators of ``ConstructGraph.py:781-782``).  The positions are the persons' peak centres of the scoremap, moved by
    up to one pixel; about ``drop`` of the joints are not annotated.  Same random stream as ``synth_scoremap``.
    """
    H = size
    W = size if width is None else width
    rng = np.random.default_rng(1000 + image_index)
    centres = rng.uniform(0.15, 0.85, size=(persons, 2)) * np.array([W, H])
    n_peaks = 2 * k + 8
    body_pts = np.zeros((num_joints, persons, 2), dtype=np.int64)
    for j in range(num_joints):                      # replay synth_scoremap's draws to recover the body peaks
        jit = rng.normal(0.0, 0.1, size=(persons, 2)) * np.array([W, H])
        body = centres + jit
        rnd = rng.uniform(0, 1, size=(n_peaks, 2)) * np.array([W, H])
        rnd[:min(persons, n_peaks)] = body[:min(persons, n_peaks)]
        body_pts[j, :, 0] = np.clip(np.rint(rnd[:persons, 0]), _R, W - 1 - _R)
        body_pts[j, :, 1] = np.clip(np.rint(rnd[:persons, 1]), _R, H - 1 - _R)
        amps = np.linspace(0.2, 0.95, n_peaks, dtype=np.float32)
        rng.shuffle(amps)
    rng2 = np.random.default_rng(3000 + image_index)
    gt = np.zeros((max_persons, num_joints, 3), dtype=np.float32)
    factors = np.ones((max_persons, num_joints), dtype=np.float32)
    for p in range(min(persons, max_persons)):
        gt[p, :, 0] = body_pts[:, p, 0] + rng2.uniform(-1.0, 1.0, size=num_joints)
        gt[p, :, 1] = body_pts[:, p, 1] + rng2.uniform(-1.0, 1.0, size=num_joints)
        gt[p, :, 2] = (rng2.uniform(size=num_joints) >= drop).astype(np.float32)
        factors[p] = rng2.uniform(20.0, 80.0, size=num_joints)
    gt[:, :, :2] *= gt[:, :, 2:3]
    return gt, factors
