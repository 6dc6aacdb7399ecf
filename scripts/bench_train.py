"""Training-step measurement (BASELINE.json configs[4]): class_agnostic_end2end shape --
agnostic MPLayer, max aggregation, skip, 10 steps -- on 8 synthetic 256x256 images per GPU (the train-time
``forward`` path works on the half-resolution maps), 30 candidates per joint, kNN-50 graph.

One step = graph construction (CUDA) -> MPN forward in train() mode (CUDA, BatchNorm batch statistics) -> a focal-free
stand-in loss (BCE-with-logits / cross-entropy on random labels: the reference's losses and label construction are
outside the hot path) -> reverse pass (CUDA) -> ONE all-reduce of the 101 203-parameter gradient bucket (NCCL) -> Adam.
Timed on the device with CUDA events, max over ranks; rank 0 prints one JSON line.

    python scripts/bench_train.py [--steps 10 --warmup 3 --batch 8]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/bench_train.py
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import pgmp_b200  # noqa: E402
import pgmp_b200._native as nv  # noqa: E402
import pgmp_b200.parallel as par  # noqa: E402
import pgmp_b200.synthetic as synthetic  # noqa: E402
from pgmp_b200.graph_constructor import get_graph_constructor  # noqa: E402
from pgmp_b200.Models.MessagePassingNetwork import get_mpn_model  # noqa: E402


def parse(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--mpn-steps", type=int, default=10)
    ap.add_argument("--model", choices=("agnostic", "flagship"), default="agnostic",
                    help="agnostic = class_agnostic_end2end (BASELINE configs[4]); flagship = the per-type / attention layer (hybrid_*)")
    ap.add_argument("--materialised-features", action="store_true",
                    help="node features gathered from a given [B,128,H,W] map (no feature_gather parameters in the step); default: "
                         "ConvUpsampleFeatures over a 32-channel backbone map, the feature_gather gradients join the all-reduce")
    ap.add_argument("--profile", action="store_true", help="after the timed steps: one more step with per-kernel CUDA events")
    return ap.parse_args(argv)


def run_training(args, rank, world, dev):
    """The measurement itself; the caller owns the process group.  Returns the result dict on rank 0 (None elsewhere)."""
    J, K = 17, 30
    data = synthetic.synth_batch(args.batch, J, args.size, K, persons=8, first_index=rank * args.batch)
    t = {k: torch.from_numpy(v).to(dev) for k, v in data.items()}
    gts, facs = zip(*[synthetic.synth_joints_gt(rank * args.batch + b, J, args.size, K, persons=8) for b in range(args.batch)])
    joints_gt, factors = torch.from_numpy(np.stack(gts)).to(dev), torch.from_numpy(np.stack(facs)).to(dev)
    # labels from the ground truth as the reference's training loop gets them (EDGE_LABEL_METHOD 6, class_agnostic_end2end)
    gcfg = pgmp_b200.config.bench_gc_config(k=K, graph_type="knn", EDGE_LABEL_METHOD=6, MATCHING_RADIUS=0.5)
    mcfg = (pgmp_b200.config.flagship_mpn_config if args.model == "flagship" else pgmp_b200.config.agnostic_mpn_config)(J, STEPS=args.mpn_steps)
    model = synthetic.synth_mpn_state_dict(get_mpn_model(mcfg), 0).to(dev).train()
    # the MPN's 101 203 (agnostic) parameters + the 36 992 of feature_gather (Conv2d(32, 128, 3, 1, 1),
    # PoseEstimation.py:64-66) are trained and all-reduced; the HRNet backbone stays out (its 32-channel map is an input)
    params = list(model.parameters())
    features = t["features"]
    if not args.materialised_features:
        from pgmp_b200.graph_constructor import ConvUpsampleFeatures
        torch.manual_seed(7)
        feature_gather = torch.nn.Conv2d(32, features.shape[1], 3, 1, 1).to(dev)
        backbone_map = torch.randn(args.batch, 32, args.size, args.size, device=dev,
                                   generator=torch.Generator(device=dev).manual_seed(100 + rank))
        params += list(feature_gather.parameters())
    opt = torch.optim.Adam(params, lr=1e-4)
    gen = torch.Generator(device=dev).manual_seed(rank)
    ev = {k: [torch.cuda.Event(enable_timing=True) for _ in range(2)] for k in ("gc", "fwd", "loss", "bwd", "ar", "opt", "step")}
    acc = {k: 0.0 for k in ev}
    per_step = {k: [] for k in ev}
    info = {}

    def step(timed):
        ev["step"][0].record()
        ev["gc"][0].record()
        feats = features if args.materialised_features else ConvUpsampleFeatures(backbone_map, feature_gather, (args.size, args.size))
        ret = get_graph_constructor(gcfg, scoremaps=t["scoremaps"], tagmaps=t["tagmaps"], features=feats,
                                    joints_gt=joints_gt, factor_list=factors, masks=None, device=dev, testing=False,
                                    heatmaps=None, num_joints=J).construct_graph()
        x, edge_attr, edge_index, joint_det = ret[0], ret[1], ret[2], ret[7]
        edge_labels, node_labels, node_classes, label_mask, node_mask, class_mask = ret[3], ret[4], ret[5], ret[8], ret[9], ret[10]
        ev["gc"][1].record()
        N, E = x.shape[0], edge_index.shape[1]
        info.update(nodes=N, edges=E)
        ev["fwd"][0].record()
        pe, pn, pc, _ = model(x, edge_attr, edge_index, node_types=joint_det[:, 2])
        ev["fwd"][1].record()
        ev["loss"][0].record()
        # masked losses on the constructor's labels (stock torch ops standing in for the reference's focal / CE losses)
        loss = (F.binary_cross_entropy_with_logits(pe[-1], edge_labels, weight=label_mask, reduction="sum") / label_mask.sum().clamp(min=1)
                + F.binary_cross_entropy_with_logits(pn[-1], node_labels, weight=node_mask, reduction="sum") / node_mask.sum().clamp(min=1)
                + (F.cross_entropy(pc[-1], node_classes, reduction="none") * class_mask).sum() / class_mask.sum().clamp(min=1))
        opt.zero_grad(set_to_none=True)
        ev["loss"][1].record()
        ev["bwd"][0].record()
        loss.backward()
        ev["bwd"][1].record()
        ev["ar"][0].record()
        info["allreduce_bytes"] = par.allreduce_gradients(params)
        info["trained_parameters"] = sum(p.numel() for p in params)
        ev["ar"][1].record()
        ev["opt"][0].record()
        opt.step()
        ev["opt"][1].record()
        ev["step"][1].record()
        torch.cuda.synchronize()
        if timed:
            for k in ev:
                acc[k] += ev[k][0].elapsed_time(ev[k][1])
                per_step[k].append(ev[k][0].elapsed_time(ev[k][1]))
        return float(loss.detach())

    for _ in range(args.warmup):
        step(False)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = nv.kernel_launches()
    losses = [step(True) for _ in range(args.steps)]
    launches = nv.kernel_launches() - launches0
    ms = {k: par.max_over_ranks(v / args.steps, dev) for k, v in acc.items()}
    # the phases are driven by the host (stock torch losses, Python between the calls): one host hiccup in a 5-step run moves
    # the mean by milliseconds, so the per-phase medians are reported next to the means
    med = {k: par.max_over_ranks(sorted(v)[len(v) // 2], dev) for k, v in per_step.items()}
    result = None
    if rank == 0:
        imgs = args.batch * world
        result = ({
            "metric": "training step: images/sec (GC + MPN forward + backward + gradient all-reduce + Adam)",
            "value": imgs / (ms["step"] * 1e-3), "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms["step"], "ms": ms, "ms_median": med, "value_median": imgs / (med["step"] * 1e-3), "edges_per_s": info["edges"] * world / (ms["step"] * 1e-3),
            "config": {"workload": "configs[4]: %d synthetic %dx%d images per GPU, %s (skip, %d steps), kNN-50 graph"
                                   % (args.batch, args.size, args.size,
                                      "per-type TypeAwareMPNLayer with attention" if args.model == "flagship" else "agnostic MPLayer (max)",
                                      args.mpn_steps),
                       "nodes_per_gpu": info["nodes"], "edges_per_gpu": info["edges"]},
            "dtype": "f32", "data": "synthetic", "scaling": "weak", "allreduce_bytes": info["allreduce_bytes"],
            "trained_parameters": info["trained_parameters"],
            "node_features": "gathered from a given map" if args.materialised_features else
                             "ConvUpsampleFeatures (feature_gather conv + interpolation at the candidates, trained)",
            "gpu_launches": launches, "loss_first_last": [losses[0], losses[-1]]})
    if args.profile and rank == 0:
        import time
        marks = []

        def mark(name):
            torch.cuda.synchronize()
            marks.append((name, time.perf_counter()))
        mark("start")
        gc = get_graph_constructor(gcfg, scoremaps=t["scoremaps"], tagmaps=t["tagmaps"], features=t["features"], joints_gt=None,
                                   factor_list=None, masks=None, device=dev, testing=False, heatmaps=None, num_joints=J)
        mark("gc ctor")
        ret = gc.construct_graph()
        mark("construct_graph")
        pe, pn, pc, _ = model(ret[0], ret[1], ret[2], node_types=ret[7][:, 2])
        mark("mpn forward")
        loss = pe[-1].sum() + pn[-1].sum() + pc[-1].sum()
        mark("loss")
        loss.backward()
        mark("backward")
        del pe, pn, pc, loss, ret, gc
        mark("free")
        print("host wall time per phase (synchronised): " + ", ".join(
            "%s %.2f ms" % (b[0], (b[1] - a[1]) * 1e3) for a, b in zip(marks[:-1], marks[1:])))
        nv.profile(True)
        step(False)
        prof = nv.profile_collect()
        nv.profile(False)
        for name, (cnt, tot) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
            print("%-28s %5d launches %9.3f ms" % (name, cnt, tot))
    return result


def main():
    args = parse()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    result = run_training(args, rank, world, dev)
    if rank == 0:
        print(json.dumps(result))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
