"""torchrun --nproc-per-node N tools/shard_bench.py : throughput of the row-sharded HBM feature table (papers100M-shaped rows,
F=128) gathered peer-to-peer over NVLink inside the gather kernel. Every rank gathers random rows; (N-1)/N of them are remote."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nts = ge.load_package()
    for kv in sys.argv[1:]:       # name=value pairs for nb_set_option (A/B of tuning knobs)
        name, value = kv.split("=")
        nts._capi.check(nts._capi.lib().nb_set_option(name.encode(), int(value)))
        if rank == 0:
            print("option", name, value)
    from sample_based_gnn_b200 import dist as nd
    cs = nts.Cuda_Stream.on_torch_stream(local)
    V, F, N = int(os.environ.get("SHARD_V", 111_059_956 // 4)), 128, int(os.environ.get("SHARD_ROWS", 400_000))          # quarter of papers100M's vertices: 14.2 GB of rows in total
    n_local = (V - rank + world - 1) // world
    # row v holds ((v * 131 + column) mod 8191): every gathered row can be checked without seeing the peer's shard
    mine = torch.empty((n_local, F), device="cuda")
    cols = torch.arange(F, device="cuda", dtype=torch.int64)[None, :]
    for a in range(0, n_local, 1 << 21):
        b = min(n_local, a + (1 << 21))
        gid = torch.arange(a, b, device="cuda", dtype=torch.int64)[:, None] * world + rank
        mine[a:b] = ((gid * 131 + cols) % 8191).float()
    st = nd.ShardedTable(cs, mine, V, F)
    g = torch.Generator(device="cuda").manual_seed(rank)
    ids = [torch.randint(0, V, (N,), device="cuda", dtype=torch.int32, generator=g) for _ in range(8)]
    out = torch.empty((N, F), device="cuda")
    for i in range(3):
        st.gather(out, ids[i], N)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 20
    for i in range(reps):
        st.gather(out, ids[i % 8], N)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    # per-peer probe: ids that all live on one shard
    probe = []
    for peer in range(world):
        pid = (torch.randint(0, V // world - 1, (N,), device="cuda", dtype=torch.int32, generator=g) * world + peer).to(torch.int32)
        st.gather(out, pid, N)
        torch.cuda.synchronize()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(5):
            st.gather(out, pid, N)
        p1.record(); torch.cuda.synchronize()
        probe.append(N * F * 4 / (p0.elapsed_time(p1) / 5) / 1e6)
    allms = [None] * world
    dist.all_gather_object(allms, (round(ms, 4), [round(x) for x in probe]))
    if rank == 0:
        for r, x in enumerate(allms):
            print(f"SHARD_RANK {r}: all-peers {x[0]} ms; per-peer GB/s {x[1]}")
    # correctness of every row of one gather (local and remote)
    st.gather(out, ids[0], N)
    torch.cuda.synchronize()
    want = ((ids[0].to(torch.int64)[:, None] * 131 + cols) % 8191).float()
    assert torch.equal(out, want), "sharded gather returned wrong rows"
    del want
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        bytes_read = N * F * 4
        print(f"SHARD_BENCH world={world} rows={N} F={F} ms={t.item():.4f} per-GPU gathered {bytes_read / t.item() / 1e6:.1f} GB/s "
              f"(remote fraction {(world - 1) / world:.3f} -> {bytes_read * (world - 1) / world / t.item() / 1e6:.1f} GB/s over NVLink per GPU); "
              f"algorithmic (4+8F)/row: {N * (4 + 8 * F) / t.item() / 1e6:.1f} GB/s")
    st.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
