"""Launched by tests/test_gpu_multi.py under torch.distributed.run: one rank per GPU. Checks the row-sharded HBM feature
table read peer-to-peer inside the gather kernel, and the bucketed NCCL gradient exchange."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nts = ge.load_package()
    from sample_based_gnn_b200 import dist as nd
    cs = nts.Cuda_Stream.on_torch_stream(local)
    V, F = 50000, 128
    full = torch.from_numpy(np.random.default_rng(1).standard_normal((V, F)).astype(np.float32))
    mine = full[rank::world].cuda()
    for pitch in (F, F + 4):
        st = nd.ShardedTable(cs, mine, V, F, pitch=pitch)
        ids = torch.from_numpy(np.random.default_rng(2 + rank).integers(0, V, 20000).astype(np.int32)).cuda()
        out = torch.empty((20000, F), device="cuda")
        st.gather(out, ids, 20000)
        torch.cuda.synchronize()
        assert torch.equal(out.cpu(), full[ids.cpu().long()]), f"rank {rank}: sharded gather mismatch (pitch {pitch})"
        st.close()
    # papers100M-shaped tiering: the hot cache (every 3rd vertex) is partitioned over the ranks and read over NVLink, the cold
    # rows are staged from this rank's host table on a side stream
    hot = np.arange(0, V, 3, dtype=np.uint32)
    hashmap = np.full(V, 0xFFFFFFFF, np.uint32)
    hashmap[hot] = np.arange(hot.size, dtype=np.uint32)
    cache = full[torch.from_numpy(hot.astype(np.int64))] + 100.0       # distinguishable from the cold copy
    st = nd.ShardedTable(cs, cache[rank::world].cuda(), hot.size, F)
    stage = nts.ColdStage(cs, full, max_rows=20000)
    d_hash = torch.from_numpy(hashmap.view(np.int32)).cuda()
    ids_np = np.random.default_rng(9 + rank).integers(0, V, 20000).astype(np.uint32)
    ids = torch.from_numpy(ids_np.view(np.int32)).cuda()
    out = torch.empty((20000, F), device="cuda")
    stage.submit(0, ids, 20000, d_hash)
    n_cold = stage.gather_table(0, out, st.table, d_hash, ids)
    torch.cuda.synchronize()
    want = full[torch.from_numpy(ids_np.astype(np.int64))].clone()
    is_hot = hashmap[ids_np] != 0xFFFFFFFF
    want[torch.from_numpy(is_hot)] += 100.0
    assert n_cold == int((~is_hot).sum()) and torch.equal(out.cpu(), want), f"rank {rank}: tiered gather mismatch"
    del stage
    st.close()
    # one-kernel all-reduce over peer memory: bit-identical to the rank-ordered fp32 sum, on every rank, for ragged sizes, many
    # back-to-back exchanges (slot / flag reuse) and ranks that arrive at different times
    par = nd.PeerAllReduce(cs, 602 * 128 + 128 * 41)
    for it, n in enumerate([1, 3, 4, 5, 1000, 602 * 128 + 128 * 41, 77, 82304] * 6):
        gens = [torch.Generator().manual_seed(1000 * it + r) for r in range(world)]
        parts = [torch.randn(n, generator=g) for g in gens]
        want = parts[0].clone()
        for q in parts[1:]:
            want += q
        mine_t = parts[rank].cuda()
        if it % 5 == rank % 5:
            torch.cuda._sleep(2_000_000)          # this rank arrives ~1 ms late
        par.all_reduce(mine_t)
        assert torch.equal(mine_t.cpu(), want), f"rank {rank}: peer all-reduce mismatch at exchange {it} (n={n})"
    # split form: push, unrelated work in the same stream, then reduce; late ranks; stats; begin/end misuse is rejected
    filler = torch.empty(1 << 22, device="cuda")
    for it in range(24):
        n = [82304, 5, 4096][it % 3]
        parts = [torch.randn(n, generator=torch.Generator().manual_seed(5000 + 10 * it + r)) for r in range(world)]
        want = parts[0].clone()
        for q in parts[1:]:
            want += q
        mine_t = parts[rank].cuda()
        if it % 4 == rank % 4:
            torch.cuda._sleep(1_000_000)
        par.begin(mine_t)
        filler.fill_(float(it))                    # the work that hides the skew
        out_t = torch.empty_like(mine_t)
        par.end(out_t)
        assert torch.equal(out_t.cpu(), want), f"rank {rank}: split exchange mismatch at {it} (n={n})"
        assert torch.equal(mine_t.cpu(), parts[rank]), "begin() must not modify its input"
    n_ex, mean_us, max_us = par.stats()
    assert n_ex >= 24 and max_us < 5e6, (n_ex, mean_us, max_us)
    t4 = torch.ones(4, device="cuda")
    par.begin(t4)
    try:
        par.begin(t4)
        raise AssertionError("a second begin() before end() must be rejected")
    except nts.NtsError:
        pass
    par.end(t4)
    assert torch.equal(t4.cpu(), torch.full((4,), float(world)))
    assert not par.timed_out()
    wp = torch.nn.Parameter(torch.zeros(64, 3, device="cuda"))
    wp.grad = torch.full_like(wp, float(rank + 1))
    nd.GradBucket([wp], peer=par).all_reduce()
    assert torch.equal(wp.grad, torch.full_like(wp, float(world * (world + 1) // 2)))
    par.close()
    w = torch.nn.Parameter(torch.zeros(602, 128, device="cuda"))
    w.grad = torch.full_like(w, float(rank + 1))
    nd.GradBucket([w]).all_reduce()
    assert torch.equal(w.grad, torch.full_like(w, float(world * (world + 1) // 2)))
    dist.barrier()
    if rank == 0:
        print(f"MGPU_OK world={world}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
