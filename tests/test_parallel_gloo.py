"""The N > 1 path on CPU: two gloo processes shard a batch of images by rank, build their shard's graph
(with the oracle -- the point here is the sharding / offset / reduction logic, not the kernels) and the
batch-ordered gather must equal the single-process graph of the whole batch."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, batch, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    import pgmp_b200
    import pgmp_b200.parallel as par
    import pgmp_b200.synthetic as synthetic

    J, S, K = 5, 64, 4
    cfg = pgmp_b200.config.bench_gc_config(k=K, graph_type="knn")
    lo, hi = par.shard_range(batch, rank, world)
    data = synthetic.synth_batch(hi - lo, J, S, K, channels=8, persons=2, first_index=lo)
    g = oracle.gc.construct_graph(data["scoremaps"], data["tagmaps"], data["features"], cfg, J)
    part = {k: torch.from_numpy(g[k]) for k in ("x", "edge_attr", "edge_index", "joint_det", "joint_scores",
                                                "batch_index", "joint_tags")}
    part["num_images"] = hi - lo
    full = par.gather_graphs(part)
    t_max = par.max_over_ranks(1.0 + rank)
    n_sum = par.sum_over_ranks(hi - lo)
    dist.barrier()
    if rank == 0:
        np.savez(os.path.join(out_dir, "gathered.npz"), t_max=t_max, n_sum=n_sum,
                 **{k: v.numpy() for k, v in full.items() if torch.is_tensor(v)})
    dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    import pgmp_b200.parallel as par
    for total in (1, 7, 32, 33):
        for world in (1, 2, 4, 8):
            cover = []
            for r in range(world):
                lo, hi = par.shard_range(total, r, world)
                cover += list(range(lo, hi))
            assert cover == list(range(total))


@pytest.mark.timeout(300)
def test_two_rank_gloo_gather_equals_single_process(tmp_path):
    import oracle
    import pgmp_b200
    import pgmp_b200.synthetic as synthetic

    batch, world, port = 5, 2, 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, batch, str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(str(tmp_path), "gathered.npz"))
    J, S, K = 5, 64, 4
    cfg = pgmp_b200.config.bench_gc_config(k=K, graph_type="knn")
    data = synthetic.synth_batch(batch, J, S, K, channels=8, persons=2)
    want = oracle.gc.construct_graph(data["scoremaps"], data["tagmaps"], data["features"], cfg, J)
    for k in ("x", "edge_attr", "edge_index", "joint_det", "joint_scores", "batch_index", "joint_tags"):
        assert np.array_equal(got[k], want[k]), k
    assert float(got["t_max"]) == 2.0 and float(got["n_sum"]) == batch


def _grad_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import pgmp_b200
    import pgmp_b200.parallel as par
    from pgmp_b200.Models.MessagePassingNetwork import get_mpn_model

    model = get_mpn_model(pgmp_b200.config.agnostic_mpn_config(17, STEPS=2))
    g = torch.Generator().manual_seed(7 + rank)
    for i, p in enumerate(model.parameters()):
        if i % 5 != 4 or rank == 0:                      # some parameters have no gradient on rank 1
            p.grad = torch.randn(p.shape, generator=g)
    nbytes = par.allreduce_gradients(model.parameters())
    if rank == 0:
        np.savez(os.path.join(out_dir, "grads.npz"), nbytes=nbytes, **{n: p.grad.numpy() for n, p in model.named_parameters()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gradient_allreduce_averages_one_bucket(tmp_path):
    """Training row: one bucket, averaged over the ranks; missing gradients count as zeros."""
    import pgmp_b200
    from pgmp_b200.Models.MessagePassingNetwork import get_mpn_model

    world, port = 2, 31000 + os.getpid() % 2000
    mp.spawn(_grad_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(str(tmp_path), "grads.npz"))
    model = get_mpn_model(pgmp_b200.config.agnostic_mpn_config(17, STEPS=2))
    total = 0
    gens = [torch.Generator().manual_seed(7 + r) for r in range(world)]
    for i, (n, p) in enumerate(model.named_parameters()):
        parts = []
        for r in range(world):
            if i % 5 != 4 or r == 0:
                parts.append(torch.randn(p.shape, generator=gens[r]))
        want = sum(parts) / world
        assert np.allclose(got[n], want.numpy(), rtol=0, atol=1e-6), n
        total += p.numel()
    assert int(got["nbytes"]) == 4 * total == 4 * 101203      # 101 203 MPN parameters
