"""world_size-2 gloo tests (CPU) of the data-parallel host logic: seed sharding covers every seed exactly once,
the bucketed gradient exchange is a SUM over ranks (core/NtsScheduler.hpp:830-836), bench's sharding matches the package's."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as ge
    nts = ge.load_package()
    from sample_based_gnn_b200 import dist as nd
    ids = np.random.default_rng(7).permutation(10007).astype(np.uint32)
    mine = nd.shard_seeds(ids, rank, world)
    inter = nd.interleave_seeds(ids[:8192], rank, world, 1024)
    w1 = torch.nn.Parameter(torch.zeros(5, 3))
    w2 = torch.nn.Parameter(torch.zeros(4))
    w1.grad = torch.full((5, 3), float(rank + 1))
    w2.grad = torch.arange(4, dtype=torch.float32) * (rank + 1)
    nd.GradBucket([w1, w2]).all_reduce()
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine.tolist(), inter.tolist()))
    if rank == 0:
        q.put((gathered, w1.grad.clone(), w2.grad.clone(), ids))
    dist.barrier()
    dist.destroy_process_group()


def test_seed_sharding_and_gradient_bucket_world2():
    world, port = 2, 29611
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, g1, g2, ids = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    shards = [np.array(g[0], dtype=np.uint32) for g in gathered]
    assert np.array_equal(np.concatenate(shards), ids)             # contiguous, complete, disjoint
    assert abs(shards[0].size - shards[1].size) <= 1
    inter = [np.array(g[1], dtype=np.uint32) for g in gathered]
    assert all(x.size == 4096 for x in inter)
    assert np.array_equal(np.sort(np.concatenate(inter)), np.sort(ids[:8192]))
    assert np.array_equal(inter[0][:512], ids[:512]) and np.array_equal(inter[1][:512], ids[512:1024])
    assert torch.equal(g1, torch.full((5, 3), 3.0))                 # 1 + 2: a sum, not a mean
    assert torch.equal(g2, torch.arange(4, dtype=torch.float32) * 3)


def _fd_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import tempfile
    import __graft_entry__ as ge
    ge.load_package()
    from sample_based_gnn_b200 import dist as nd
    with tempfile.TemporaryFile() as f:           # stands in for the exported shard descriptor
        f.write(f"shard-of-rank-{rank}".encode())
        f.flush()
        got = nd.exchange_fds(f.fileno())
        seen = {}
        for r, fd in got.items():
            seen[r] = os.pread(fd, 64, 0).decode()
            os.close(fd)
    ok = sorted(seen) == [r for r in range(world) if r != rank] and all(v == f"shard-of-rank-{r}" for r, v in seen.items())
    oks = [None] * world
    dist.all_gather_object(oks, ok)
    if rank == 0:
        q.put(oks)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_descriptor_exchange_world3():
    """The sharded table hands each shard's POSIX descriptor to every other rank (SCM_RIGHTS over unix sockets)."""
    world, port = 3, 29613
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_fd_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    oks = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert oks == [True] * world


def test_bench_sharding_matches_package():
    sys.path.insert(0, ROOT)
    import bench
    ids = bench.train_seeds(5000)
    parts = [bench.shard_seeds(ids, r, 4) for r in range(4)]
    assert np.array_equal(np.concatenate(parts), ids)
