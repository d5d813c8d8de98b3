"""torchrun --nproc-per-node N tools/shard_diag.py : why is the all-peers sharded gather slow? (a) rank 0 alone, touching k distinct
peers; (b) every rank reads only from its ring neighbour; (c) every rank reads from all peers."""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); os.environ["NCCL_DEBUG"] = "WARN"
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nts = ge.load_package()
from sample_based_gnn_b200 import dist as nd
cs = nts.Cuda_Stream.on_torch_stream(local)
V, F, N = 111_059_956 // 4, 128, 400_000
n_local = (V - rank + world - 1) // world
st = nd.ShardedTable(cs, torch.rand((n_local, F), device="cuda"), V, F)
g = torch.Generator(device="cuda").manual_seed(rank)
out = torch.empty((N, F), device="cuda")
rows = V // world - 1

def ids_for(peers):
    p = torch.tensor(peers, device="cuda", dtype=torch.int32)
    which = p[torch.randint(0, len(peers), (N,), device="cuda", generator=g)]
    return (torch.randint(0, rows, (N,), device="cuda", dtype=torch.int32, generator=g) * world + which).to(torch.int32)

def run(ids, active=True):
    dist.barrier(); torch.cuda.synchronize()
    ms = 0.0
    if active:
        for _ in range(2): st.gather(out, ids, N)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): st.gather(out, ids, N)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
    dist.barrier()
    return ms

res = {}
for k in range(1, world):
    peers = [(rank + 1 + j) % world for j in range(k)]
    res[f"solo_{k}peers"] = run(ids_for(peers), active=(rank == 0))
res["ring_all_ranks"] = run(ids_for([(rank + 1) % world]))
res["allpeers_all_ranks"] = run(ids_for(list(range(world))))
res["allpeers_solo"] = run(ids_for(list(range(world))), active=(rank == 0))
allr = [None] * world
dist.all_gather_object(allr, res)
if rank == 0:
    for k in res:
        vals = [r[k] for r in allr if r[k] > 0]
        print(f"DIAG {k:22s} ms: {[round(v, 3) for v in vals]}  -> {N * F * 4 / max(vals) / 1e6:.0f} GB/s per reader (slowest)")
st.close(); dist.destroy_process_group()
