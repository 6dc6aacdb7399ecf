"""Host-side pieces of bench.py that must never take the measurement down: the NUMA binding of multi-GPU ranks and the
clock sampler (in-process NVML, `nvidia-smi` as the fallback) degrade to "unavailable" on a machine without a GPU."""
import importlib.util
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("pgmp_bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


def test_numa_binding_is_best_effort():
    before = os.sched_getaffinity(0)
    out = bench.bind_to_gpu_numa_node(torch, 0)
    assert out is None or {"gpu", "node", "cpus"} <= set(out)
    if out is None:
        assert os.sched_getaffinity(0) == before


def test_clock_sampler_without_a_gpu_reports_unavailable():
    s = bench.ClockSampler(0, "00000000:00:00.0")
    s.start()
    out = s.summary()
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    if out["sm_mhz"] is None:
        assert out["reasons"] == ["unavailable"]


def test_workload_config_names_the_baseline_config():
    import argparse
    ns = argparse.Namespace(batch=32, precision="tc")
    cfg = bench.workload_config(ns, 8)
    assert cfg["global_batch"] == 256 and "configs[1]" in cfg["workload"] and cfg["parallelism"].startswith("dp8")
