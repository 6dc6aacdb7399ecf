#!/usr/bin/env python
"""Benchmark of the post-backbone grouping path (graph constructor + message-passing network).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload = BASELINE.json configs[1]: a COCO-shaped batch of 32 synthetic 512x512 images per GPU
(17 joints, 128-channel w32 feature map, 30 candidates per joint, symmetric kNN-50 candidate graph,
flagship per-type / edge-attention MPN with skip connections, 10 steps).  One step = one pass of
construct_graph() + mpn.forward() over the batch.  Prints ONE JSON line (rank 0).

  value        images/s with the inputs resident in HBM (device-timed with CUDA events, max over ranks), K batches
               through pgmp_b200.pipeline.GroupingPipeline (construct_graph + forward per batch; the next batch's
               detection half runs on a side stream so the host read of the counts never drains the GPU);
               serial_calls = the same batches as plain back-to-back calls
  e2e          the same metric through the same API called with HOST (pinned) tensors, every step: the heatmaps
               are copied to the device (side stream), the feature / tag maps are read in place over PCIe at the
               candidate pixels only (the constructor leaves pinned maps on the host), the logits are read back;
               stage_ms times the stages alone; all_inputs_copied copies every input first as the reference does
  roofline     dominant kernel: algorithmic bytes per launch / its mean CUDA-event duration vs measured HBM peak
  cpu_baseline the UNMODIFIED reference files (baseline/_ref, a verbatim git-ignored copy made by oracle/make_ref.py;
               kind "reference") timed on this box's host cores; the numpy oracle port (kind "port") only when no
               copy of the reference is reachable
  --impl reference: the same CPU arm as its own bench line, a bounded sample of images per step
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec and graph edges/sec (grouping path, 512px) at 1/2/4/8 B200"
J, SIZE, CAND, CHANNELS = 17, 512, 30, 128


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU per step")
    ap.add_argument("--precision", default=os.environ.get("PGMP_PRECISION", "tc"), choices=["fp32", "tc"])
    ap.add_argument("--ref-images", type=int, default=2, help="images per step of the CPU reference arm")
    ap.add_argument("--cpu-sample", type=int, default=12, help="images of the cpu_baseline sample")
    return ap.parse_args()


def workload_config(args, world):
    return {"workload": "configs[1]: %d synthetic 512x512 images per GPU, 17 joints, 128-ch w32 features, "
                        "30 candidates/joint, kNN-50 graph, flagship per-type/attention MPN (skip, 10 steps)" % args.batch,
            "images_per_gpu": args.batch, "global_batch": args.batch * world, "parallelism": "dp%d (images sharded, "
            "no collective in inference)" % world, "precision_mode": args.precision,
            "l2": "inputs (5.4 GB per batch) exceed the 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------ CPU arms
def host_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [i.get("num_threads", 1) for i in threadpool_info()]
        return max(n) if n else 1
    except Exception:
        return os.cpu_count() or 1


_CPU_ARM = {}


def cpu_arm():
    """The CPU implementation the baseline legs time: the UNMODIFIED reference files (graph constructor + MPN, loaded
    byte for byte by oracle/ref_shims.py from /root/reference or from the verbatim copy oracle/make_ref.py leaves in
    baseline/_ref) when they are reachable -- kind "reference" -- else the numpy oracle port -- kind "port"."""
    if _CPU_ARM:
        return _CPU_ARM
    import torch

    import pgmp_b200
    import pgmp_b200.synthetic as synthetic
    from oracle import ref_shims
    from pgmp_b200.Models.MessagePassingNetwork import get_mpn_model

    gcfg = pgmp_b200.config.bench_gc_config(k=CAND, graph_type="knn")
    mcfg = pgmp_b200.config.flagship_mpn_config(J)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    src = ref_shims.ref_src()
    if src is not None:
        cg, mpn = ref_shims.load_reference()
        model = synthetic.synth_mpn_state_dict(mpn.NodeClassificationMPNSimple(mcfg), 1).eval()

        def run(data):
            t = {k: torch.from_numpy(v) for k, v in data.items()}
            with torch.no_grad():
                ret = cg.NaiveGraphConstructor(
                    scoremaps=t["scoremaps"], tagmaps=t["tagmaps"], features=t["features"], joints_gt=None, factor_list=None,
                    masks=None, device=torch.device("cpu"), config=gcfg, testing=True, heatmaps=None,
                    num_joints=J).construct_graph()
                model(ret[0], ret[1], ret[2], node_labels=None, edge_labels=None, batch_index=ret[12], node_mask=None,
                      node_types=ret[7][:, 2])
            return int(ret[2].shape[1])
        _CPU_ARM.update(kind="reference", run=run, cores=threads,
                        what="the unmodified reference files (ConstructGraph.py + NodeClassificationMPNSimple.py) from %s; "
                             "torch_geometric / torch_scatter / torch_cluster are the pure-torch stand-ins of "
                             "oracle/ref_shims.py" % ("baseline/_ref" if "baseline" in src else src))
    else:
        import oracle
        model = synthetic.synth_mpn_state_dict(get_mpn_model(mcfg), 1).eval()
        sd = {k: v.numpy() for k, v in model.state_dict().items()}

        def run(data):
            g = oracle.gc.construct_graph(data["scoremaps"], data["tagmaps"], data["features"], gcfg, J)
            oracle.mpn.node_classification_mpn_forward(sd, mcfg, g["x"], g["edge_attr"], g["edge_index"], g["joint_det"][:, 2])
            return int(g["edge_index"].shape[1])
        _CPU_ARM.update(kind="port", run=run, cores=host_threads(),
                        what="numpy oracle port of the reference's CPU path (no copy of the reference is reachable: "
                             "run oracle/make_ref.py where /root/reference exists)")
    return _CPU_ARM


def cpu_images_per_s(num_images, first_index=0, warm=False):
    """Time the CPU arm (GC + flagship MPN) on `num_images` images of the workload, one at a time like the
    reference's batch-1 evaluation loop (valid.py:95); input synthesis is outside the clock.  Returns (seconds, edges)."""
    import pgmp_b200.synthetic as synthetic

    arm = cpu_arm()
    total, edges = 0.0, 0
    for i in range(num_images + (1 if warm else 0)):
        data = synthetic.synth_batch(1, J, SIZE, CAND, channels=CHANNELS, first_index=first_index + i)
        t0 = time.perf_counter()
        e = arm["run"](data)
        dt = time.perf_counter() - t0
        if warm and i == 0:
            continue
        total += dt
        edges += e
    return total, edges


def run_reference(args, rank, world):
    if rank != 0:
        return
    arm = cpu_arm()
    for _ in range(args.warmup):
        cpu_images_per_s(1)
    t, edges = 0.0, 0
    for s in range(args.steps):
        dt, e = cpu_images_per_s(args.ref_images, first_index=s * args.ref_images)
        t += dt
        edges += e
    imgs = args.steps * args.ref_images
    v = imgs / t
    cfg = workload_config(args, world)
    cfg["reference_sample"] = ("each step times %d images of the workload (a bounded sample of the %d-image batch, so that the "
                               "run ends within minutes); images/s does not depend on the sample size: the reference "
                               "loops over images (ConstructGraph.py:58)" % (args.ref_images, args.batch))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "images/s", "edges_per_s": edges / t,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": v, "unit": "images/s", "cores": arm["cores"], "kind": arm["kind"],
                             "sample": "%d images per step x %d steps of the workload, one image at a time; %s"
                                       % (args.ref_images, args.steps, arm["what"])},
            "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """SM clock and throttle reasons of this rank's GPU, sampled DURING the timed region.  In-process NVML calls
    (nvidia_ml_py, ~50 us each, every 20 ms): spawning `nvidia-smi` instead forks this process -- gigabytes of pinned
    mappings -- inside the timed region and holds the interpreter lock while it does (measured at 2 GPUs: the pipelined
    loop, whose host thread must keep launching, ran 4.7 ms per step with the fork in it against 4.0 ms for the serial
    loop without).  `nvidia-smi` remains the fallback when NVML cannot be loaded."""

    def __init__(self, index, pci_bus_id=None):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = (pynvml.nvmlDeviceGetHandleByPciBusId(pci_bus_id.encode()) if pci_bus_id
                           else pynvml.nvmlDeviceGetHandleByIndex(index))
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def run(self):
        if self.nvml is not None:
            nv_ = self.nvml
            bits = (0x8, 0x40, 0x20, 0x4)        # hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap
            while not self.stop_flag.is_set():
                try:
                    sm = nv_.nvmlDeviceGetClockInfo(self.handle, nv_.NVML_CLOCK_SM)
                    try:
                        why = nv_.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                    except Exception:
                        why = nv_.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                    self.rows.append([str(sm), str(self.max_sm)] + ["Active" if why & b else "Not Active" for b in bits])
                except Exception:
                    pass
                self.stop_flag.wait(0.02)
            return
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows), "source": "nvml (in process)" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------ GPU arm
def bind_to_gpu_numa_node(torch, local_rank):
    """Multi-GPU runs: keep this rank's threads -- and therefore the pinned host buffers it allocates next (local
    allocation policy) -- on the NUMA node its GPU hangs off, so that the 570 MB heatmap copy of every step does not cross
    the socket interconnect (round 1: end-to-end efficiency 0.65 at 8 GPUs with unbound ranks).  Best effort: returns a
    short description, or None when the topology cannot be read."""
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if len(cpus) < 2:
            return None
        os.sched_setaffinity(0, cpus)
        return {"gpu": bdf, "node": node, "cpus": len(cpus)}
    except Exception:
        return None


def run_b200(args, rank, local_rank, world):
    import numpy as np
    import torch
    import torch.distributed as dist

    import pgmp_b200
    import pgmp_b200._native as nv
    import pgmp_b200.synthetic as synthetic
    from pgmp_b200.graph_constructor import get_graph_constructor
    from pgmp_b200.Models.MessagePassingNetwork import get_mpn_model
    from pgmp_b200.Utils import group_persons

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    first = rank * B                                           # images are sharded: rank r owns images [rB, rB+B)
    sm_h = torch.from_numpy(np.stack([synthetic.synth_scoremap(first + b, J, SIZE, CAND) for b in range(B)])).pin_memory()
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    feat = torch.randn(B, CHANNELS, SIZE, SIZE, device=dev, generator=gen)
    tags = torch.randn(B, J, SIZE, SIZE, device=dev, generator=gen)
    sm = sm_h.to(dev)
    gcfg = pgmp_b200.config.bench_gc_config(k=CAND, graph_type="knn")
    mcfg = pgmp_b200.config.flagship_mpn_config(J, B200_PRECISION=args.precision)
    model = synthetic.synth_mpn_state_dict(get_mpn_model(mcfg), 1).eval().to(dev)

    def step(scoremaps, tagmaps, features):
        gc = get_graph_constructor(gcfg, scoremaps=scoremaps, tagmaps=tagmaps, features=features, joints_gt=None,
                                   factor_list=None, masks=None, device=dev, testing=True, heatmaps=None,
                                   num_joints=J)
        ret = gc.construct_graph()
        with torch.no_grad():
            pe, pn, pc, _ = model(ret[0], ret[1], ret[2], node_labels=None, edge_labels=None, batch_index=ret[12],
                                  node_mask=None, node_types=ret[7][:, 2])
        return ret, pe, pn, pc

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    from pgmp_b200.pipeline import GroupingPipeline
    pipe = GroupingPipeline(gcfg, model, J, dev)
    resident = dict(scoremaps=sm, tagmaps=tags, features=feat)

    def run_pipelined(batch, steps, after=None):
        """`steps` batches through the public pipelined API (construct_graph + forward per batch, the detection half of
        the next batch in flight on a side stream); every batch does the full work."""
        out = None
        for out in pipe.run(batch for _ in range(steps)):
            if after is not None:
                after(out)
        return out

    for _ in range(max(args.warmup, 3)):
        ret, pe, pn, pc = step(sm, tags, feat)
    run_pipelined(resident, 2)
    edges_per_step = int(ret[2].shape[1])
    nodes_per_step = int(ret[0].shape[0])

    # ---- timed region 1: inputs resident in HBM, K batches through the pipelined API
    pr_ = torch.cuda.get_device_properties(local_rank)
    sampler = ClockSampler(local_rank, "%08X:%02X:%02X.0" % (pr_.pci_domain_id, pr_.pci_bus_id, pr_.pci_device_id))
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches0 = nv.kernel_launches()
    ev0.record()
    run_pipelined(resident, args.steps)
    ev1.record()
    barrier()
    launches = nv.kernel_launches() - launches0
    t_dev = reduce_max(ev0.elapsed_time(ev1) / 1e3)
    clocks = sampler.summary()
    # the same K steps as plain back-to-back calls (one host wait per batch), for comparison
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step(sm, tags, feat)
    ev1.record()
    barrier()
    t_serial = reduce_max(ev0.elapsed_time(ev1) / 1e3)

    # ---- per-kernel CUDA-event durations over K steps (separate pass; events bracket every launch)
    nv.profile(True)
    for _ in range(args.steps):
        step(sm, tags, feat)
    prof = nv.profile_collect()
    nv.profile(False)

    # ---- timed region 2: end to end from pinned host buffers through the same API.  The heatmaps are copied to the
    # device (every pixel is read by the NMS); of the feature / tag maps only the candidate pixels are read, in place
    # over PCIe (graph_constructor/__init__.py); the logits are read back into pinned host memory every step.
    # The feature maps sit in pinned host memory in channels_last memory format ([B, C, H, W] tensor, C innermost): a
    # node's 128 channels are one 512-byte read over PCIe.  With the reference's NCHW strides the same gather is 2.1 M
    # separate 4-byte reads (measured: 7.1 ms per step, features_nchw below); both layouts give bit-identical outputs.
    feat_h = torch.empty(feat.shape, dtype=feat.dtype, pin_memory=True, memory_format=torch.channels_last).copy_(feat)
    feat_h_nchw = torch.empty(feat.shape, dtype=feat.dtype, pin_memory=True).copy_(feat) if world == 1 else None
    tags_h = torch.empty(tags.shape, dtype=tags.dtype, pin_memory=True).copy_(tags)
    host = dict(scoremaps=sm_h, tagmaps=tags_h, features=feat_h)
    out_h = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in (pe[-1], pn[-1], pc[-1])]
    d2h = sum(h.numel() * 4 for h in out_h)
    n_nodes = nodes_per_step

    def read_back(out):
        for h, d in zip(out_h, (out[1][0][-1], out[1][1][-1], out[1][2][-1])):
            h.copy_(d, non_blocking=True)

    def time_e2e(fn):
        fn(max(args.warmup, 3))                     # W >= 3 untimed steps: the first host-buffer batches also grow the
        barrier()                                   # side stream's allocator pool (one 570 MB heatmap buffer per batch in flight)
        ev0.record()
        fn(args.steps)
        ev1.record()
        barrier()
        return reduce_max(ev0.elapsed_time(ev1) / 1e3)

    def e2e_serial(steps, copy_all=False):          # plain calls: copy, compute and read-back strictly one after another
        for _ in range(steps):
            b = dict(host)
            if copy_all:                            # the reference's behaviour: every input moved first (CG.py:12-18)
                b = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
            gc = get_graph_constructor(gcfg, scoremaps=b["scoremaps"], tagmaps=b["tagmaps"], features=b["features"],
                                       joints_gt=None, factor_list=None, masks=None, device=dev, testing=True, heatmaps=None,
                                       num_joints=J)
            ret_ = gc.construct_graph()
            with torch.no_grad():
                o = model(ret_[0], ret_[1], ret_[2], node_types=ret_[7][:, 2])
            read_back((ret_, o))

    t_e2e = time_e2e(lambda k: run_pipelined(host, k, after=read_back))
    t_e2e_serial = time_e2e(e2e_serial)
    t_e2e_nchw = None
    if feat_h_nchw is not None:                     # (N = 1 only: another 4.3 GB of pinned memory per rank)
        host_nchw = dict(host, features=feat_h_nchw)
        t_e2e_nchw = time_e2e(lambda k: run_pipelined(host_nchw, k, after=read_back))
        del host_nchw, feat_h_nchw
    t_e2e_copy = time_e2e(lambda k: e2e_serial(k, copy_all=True))
    # stages of one end-to-end step, each timed alone: the heatmap copy, and the kernels with the maps on the host
    sm_d = torch.empty(sm_h.shape, dtype=sm_h.dtype, device=dev)
    sm_d.copy_(sm_h, non_blocking=True)
    barrier()
    ev0.record()
    for _ in range(3):
        sm_d.copy_(sm_h, non_blocking=True)
    ev1.record()
    barrier()
    t_copy = ev0.elapsed_time(ev1) / 3
    nv.profile(True)
    e2e_serial(2)
    prof_h = nv.profile_collect()
    nv.profile(False)
    stage_ms = {"heatmap_h2d": t_copy,
                "gather_from_host": sum(prof_h[k][1] / prof_h[k][0] for k in ("gather_features_kernel", "emit_nodes_kernel") if k in prof_h),
                "kernels_total": sum(v[1] for v in prof_h.values()) / 2}
    del sm_d
    h2d_copy_all = sm_h.numel() * 4 + feat_h.numel() * 4 + tags_h.numel() * 4
    # bytes that cross PCIe towards the device per step: the heatmap copy + the gathered feature / tag elements
    h2d = sm_h.numel() * 4 + n_nodes * feat_h.shape[1] * 4 + n_nodes * 4

    # ---- the grouping tail (sigmoid / threshold / GAEC / persons): alone on the last logits, and inside the pipelined API
    #      (batch i's grouping on a third stream while batch i + 1 goes through the network), resident and end to end
    gcx = get_graph_constructor(gcfg, scoremaps=sm, tagmaps=tags, features=feat, joints_gt=None, factor_list=None, masks=None,
                                device=dev, testing=True, heatmaps=None, num_joints=J)
    ret = gcx.construct_graph()
    with torch.no_grad():
        pe, pn, pc, _ = model(ret[0], ret[1], ret[2], node_types=ret[7][:, 2])
    gkw = dict(node_threshold=0.1, detector_scores=ret[11], nodes_per_image=gcx.num_nodes_per_image,
               edges_per_image=gcx.num_edges_per_image)
    group_persons(ret[7], pn[-1], ret[2], pe[-1], pc[-1], ret[12], J, **gkw)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        groups = group_persons(ret[7], pn[-1], ret[2], pe[-1], pc[-1], ret[12], J, **gkw)
    ev1.record()
    barrier()
    t_group = reduce_max(ev0.elapsed_time(ev1) / 1e3)
    persons_per_image = sum(0 if g_ is None else len(g_[0]) for g_ in groups) / max(len(groups), 1)
    pipe_g = GroupingPipeline(gcfg, model, J, dev, group=dict(node_threshold=0.1))

    def run_grouped(batch, steps, after=None):
        for out in pipe_g.run(batch for _ in range(steps)):
            if after is not None:
                after(out)

    run_grouped(resident, 2)
    barrier()
    ev0.record()
    run_grouped(resident, args.steps)
    ev1.record()
    barrier()
    t_dev_g = reduce_max(ev0.elapsed_time(ev1) / 1e3)
    t_e2e_g = time_e2e(lambda k: run_grouped(host, k, after=read_back))

    total_images = B * world * args.steps
    total_edges = reduce_sum(edges_per_step) * args.steps
    value = total_images / t_dev
    # ---- SURVEY.md 8f rank 2, beside the headline (N = 1): the head's two output stages assembled inside the NMS load
    #      stage (HeadStages -> pgmp_gc_detect_fused) against assemble_kernel + plain detection, same maps
    assembly = None
    if world == 1:
        try:
            from pgmp_b200.graph_constructor import HeadStages, hr_process_output
            gen2 = torch.Generator(device=dev).manual_seed(5)
            s1 = torch.cat([torch.nn.functional.avg_pool2d(sm, 2) + 0.01 * torch.rand(B, J, SIZE // 2, SIZE // 2, device=dev, generator=gen2),
                            torch.randn(B, J, SIZE // 2, SIZE // 2, device=dev, generator=gen2)], 1).contiguous()
            small = torch.zeros(B, 4, SIZE, SIZE, device=dev)

            def gc_of(scoremaps, tagmaps):
                return get_graph_constructor(gcfg, scoremaps=scoremaps, tagmaps=tagmaps, features=small, joints_gt=None,
                                             factor_list=None, masks=None, device=dev, testing=True, heatmaps=None,
                                             num_joints=J).construct_graph()

            def two_kernels():
                score, _, tg = hr_process_output(((s1, sm), None), "avg", J)
                return gc_of(score, tg)

            def fused():
                st = HeadStages((s1, sm), J)
                return gc_of(st, st)

            assembly = {}
            for nm, fn in (("assemble_then_detect", two_kernels), ("fused", fused)):
                for _ in range(3):
                    fn()
                nv.profile(True)
                for _ in range(args.steps):
                    fn()
                pa = nv.profile_collect()
                nv.profile(False)
                assembly[nm] = {k: v[1] / args.steps for k, v in pa.items()
                                if k.startswith("nms_candidates") or k.startswith("assemble") or k.startswith("stage_tags")}
                assembly[nm]["ms"] = sum(assembly[nm].values())
            # the whole path from the head's two stages (assembly + GC + MPN per batch through GroupingPipeline): the stages
            # assembled by hr_process_output in front of the constructor, against HeadStages passed straight in
            def whole_path(make_batch):
                def run(k):
                    for _ in pipe.run(make_batch() for _ in range(k)):
                        pass
                run(max(args.warmup, 3))
                barrier()
                ev0.record()
                run(args.steps)
                ev1.record()
                barrier()
                return ev0.elapsed_time(ev1) / args.steps

            def batch_two():
                score, _, tg = hr_process_output(((s1, sm), None), "avg", J)
                return dict(scoremaps=score, tagmaps=tg, features=feat)

            def batch_fused():
                st = HeadStages((s1, sm), J)
                return dict(scoremaps=st, tagmaps=st, features=feat)

            assembly["whole_path_ms_per_step"] = {"assemble_then_pipeline": whole_path(batch_two), "head_stages_fused": whole_path(batch_fused)}
            rd = B * J * (SIZE * SIZE + (SIZE // 2) ** 2) * 4
            assembly["fused_bytes_read"] = rd
            assembly["fused_gbs"] = rd / assembly["fused"]["ms"] / 1e6
            assembly["note"] = ("hr_process_output (hrnet.py:587-611) + NMS: the fused kernel reads both stages once and writes "
                                "no map; the two-kernel path writes and re-reads the 570 MB map and the up-sampled tag maps")
            del s1, small
        except Exception as exc:
            assembly = {"error": "%s: %s" % (type(exc).__name__, exc)}
    # BASELINE.json configs[4] beside the headline: one training step of the agnostic MPN
    # (GC + forward + reverse pass + gradient all-reduce + Adam), same ranks; scripts/bench_train.py is the stand-alone form
    train = None
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("pgmp_bench_train", os.path.join(ROOT, "scripts", "bench_train.py"))
        bt = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(bt)
        del feat, tags, sm
        torch.cuda.empty_cache()
        r = bt.run_training(bt.parse(["--steps", "10", "--warmup", "3"]), rank, world, dev)
        keys = ("value", "unit", "ms_per_step", "ms", "ms_median", "value_median", "edges_per_s", "config", "allreduce_bytes",
                "trained_parameters", "node_features", "gpu_launches")
        if rank == 0:
            train = {k: r[k] for k in keys}
        # the same step with the flagship per-type / attention layer (TypeAwareMPNLayer, the hybrid_* configs)
        r = bt.run_training(bt.parse(["--steps", "5", "--warmup", "3", "--model", "flagship"]), rank, world, dev)
        if rank == 0:
            train["flagship"] = {k: r[k] for k in keys}
        # ... and the agnostic step with its E-level forward products on tcgen05 (PGMP_TRAIN_TC=1, csrc/mpn_train_tc.cu:
        # bf16 hi / lo operand pairs, 1e-5-level forward; the rows above use the 3xTF32 kernels the parity tests pin)
        os.environ["PGMP_TRAIN_TC"] = "1"
        try:
            r = bt.run_training(bt.parse(["--steps", "5", "--warmup", "3"]), rank, world, dev)
        finally:
            os.environ.pop("PGMP_TRAIN_TC", None)
        if rank == 0:
            train["tc_forward"] = {k: r[k] for k in keys}
    except Exception as exc:      # the headline line must not depend on the training row
        train = dict(train or {}, error="%s: %s" % (type(exc).__name__, exc))
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (largest total CUDA-event time)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    top = max(prof.items(), key=lambda kv: kv[1][1])
    name, (cnt, ms) = top
    # algorithmic bytes per launch ("Kernels"): SURVEY.md 8(d) per-unit figure x units per launch
    elem = 4                                                     # storage width of edge features in this mode
    bytes_per_launch = {
        "edge_step_kernel": 3 * 64 * elem * edges_per_step,       # read g, read C (skip constant), write g'
        "edge_step_tc_kernel": 3 * 64 * elem * edges_per_step,     # g and g' are bf16 hi+lo = 4 bytes per element
    }.get(name)
    if bytes_per_launch is None and name.startswith("nms_candidates"):
        bytes_per_launch = J * SIZE * SIZE * 4 * B + nodes_per_step * 28
    if bytes_per_launch is None:
        # a kernel without a byte model became the largest one: report the message-passing step kernel (the kernel the
        # roofline is defined for) instead of dropping the key, and say so
        for cand in ("edge_step_tc_kernel", "edge_step_kernel"):
            if cand in prof:
                name, (cnt, ms) = cand, prof[cand]
                bytes_per_launch = 3 * 64 * elem * edges_per_step
                break
    roofline = None
    traffic = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture
        tj = {}
        for rnd in ("r1", "r2"):                      # the latest committed capture wins
            fn = os.path.join(ROOT, "profiles", rnd + "_traffic.json")
            if os.path.exists(fn):
                tj.update(json.load(open(fn)))
        ent = tj.get(name)
        if ent and ent.get("edges_per_launch") == edges_per_step:
            traffic = ent["dram_bytes_per_launch"]
    except Exception:
        pass
    if bytes_per_launch:
        dur = ms / cnt / 1e3
        ach = bytes_per_launch / dur / 1e9
        roofline = {"bound": "hbm", "kernel": name, "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ach / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                    "launch_ms": 1e3 * dur, "launches": cnt,
                    "share_of_kernel_time": ms / sum(v[1] for v in prof.values()),
                    "algorithmic_bytes_per_launch": bytes_per_launch}
    kernels = {k: {"launches": c, "ms": round(m, 4)} for k, (c, m) in sorted(prof.items(), key=lambda kv: -kv[1][1])[:8]}
    nms = next(((k, v) for k, v in prof.items() if k.startswith("nms_candidates")), None)
    roofline_nms = None
    if nms:
        nb = J * SIZE * SIZE * 4 * B
        nd_ = nms[1][1] / nms[1][0] / 1e3
        roofline_nms = {"bound": "hbm", "kernel": nms[0], "achieved": nb / nd_ / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": nb / nd_ / 1e9 / hbm_peak, "launch_ms": 1e3 * nd_, "algorithmic_bytes_per_launch": nb}

    cpu_t, cpu_edges = (None, None)
    cpu = None
    if world == 1:
        cpu_t, cpu_edges = cpu_images_per_s(args.cpu_sample, warm=True)
        arm = cpu_arm()
        cpu = {"value": args.cpu_sample / cpu_t, "unit": "images/s", "cores": arm["cores"], "kind": arm["kind"],
               "edges_per_s": cpu_edges / cpu_t,
               "sample": "%d images of the workload, one at a time; %s" % (args.cpu_sample, arm["what"])}

    line = {"metric": METRIC, "value": value, "unit": "images/s", "edges_per_s": total_edges / t_dev,
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * t_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16x3 (bf16 hi/lo operand pairs on tcgen05, fp32 accumulate, fp32-equivalent storage)" if args.precision == "tc" else "f32",
            "data": "synthetic", "config": workload_config(args, world), "clocks": clocks,
            "serial_calls": {"value": total_images / t_serial, "ms_per_step": 1e3 * t_serial / args.steps,
                             "note": "the same batches as plain back-to-back construct_graph() + forward() calls (one host "
                                     "wait per batch); value uses pgmp_b200.pipeline.GroupingPipeline, which starts the "
                                     "next batch's detection half on a side stream"},
            "e2e": {"value": total_images / t_e2e, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * t_e2e / args.steps,
                    "inputs": "pinned host tensors passed to the public API (GroupingPipeline): heatmaps copied on the side "
                              "stream while the previous batch computes, feature / tag maps gathered in place over PCIe "
                              "(only the candidate pixels), logits read back every step",
                    "stage_ms": stage_ms,
                    "pcie_gbs_heatmap_copy": sm_h.numel() * 4 / (t_copy * 1e-3) / 1e9,
                    "serial_calls": {"value": total_images / t_e2e_serial, "ms_per_step": 1e3 * t_e2e_serial / args.steps},
                    "features_nchw": None if t_e2e_nchw is None else {
                        "value": total_images / t_e2e_nchw, "ms_per_step": 1e3 * t_e2e_nchw / args.steps,
                        "note": "the same pipelined calls with the host feature maps in NCHW strides"},
                    "all_inputs_copied": {"value": total_images / t_e2e_copy, "ms_per_step": 1e3 * t_e2e_copy / args.steps,
                                          "h2d_bytes_per_step": h2d_copy_all}},
            "gpu_launches": int(launches), "roofline": roofline, "roofline_nms": roofline_nms, "cpu_baseline": cpu,
            "kernels": kernels,
            "grouping_tail": {"ms_per_step": 1e3 * t_group / args.steps, "persons_per_image": persons_per_image,
                              "note": "sigmoid / threshold / GAEC / persons on the step's logits, alone (not part of value: "
                                      "the metric is GC + MPN, SURVEY.md 8d)",
                              "with_grouping": {"value": B * world * args.steps / t_dev_g, "ms_per_step": 1e3 * t_dev_g / args.steps,
                                                "e2e_value": B * world * args.steps / t_e2e_g,
                                                "e2e_ms_per_step": 1e3 * t_e2e_g / args.steps,
                                                "note": "GC + MPN + grouping per batch through GroupingPipeline(group=...): "
                                                        "the grouping of batch i overlaps the network of batch i + 1; persons "
                                                        "copied to the host every step"}},
            "graph": {"nodes_per_step_per_gpu": nodes_per_step, "edges_per_step_per_gpu": edges_per_step},
            "numa_binding_rank0": numa}
    line["train_step"] = train
    line["scoremap_assembly"] = assembly
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
