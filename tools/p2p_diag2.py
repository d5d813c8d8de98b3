"""single process, all GPUs visible: peer-read gather through plain (non-IPC) peer pointers"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
nts = ge.load_package()
n = torch.cuda.device_count()
print("gpus", n)
V, F, N = 111_059_956 // 4, 128, 400_000
rows = V // n
shards = []
for d in range(n):
    shards.append(torch.rand((rows + 1, F), device=f"cuda:{d}"))
# torch enables peer access on first p2p copy
for d in range(1, n):
    _ = shards[d][:4].to("cuda:0")   # a p2p copy makes torch enable peer access 0 <- d
torch.cuda.synchronize(0)
a = torch.empty(64 * 1024 * 1024, device="cuda:0"); b = torch.empty_like(a, device="cuda:1")
for _ in range(2): a.copy_(b)
torch.cuda.synchronize(0); t = time.time()
for _ in range(10): a.copy_(b)
torch.cuda.synchronize(0); torch.cuda.synchronize(1)
print("torch peer copy 1->0 GB/s", 10 * a.numel() * 4 / (time.time() - t) / 1e9)
torch.cuda.set_device(0)
cs = nts.Cuda_Stream.on_torch_stream(0)
table = nts.FeatureTable(cs, [s.data_ptr() for s in shards], F, F, rows * n, keepalive=shards)
out = torch.empty((N, F), device="cuda:0")
g = torch.Generator(device="cuda:0").manual_seed(0)
for k in range(1, n + 1):
    peers = torch.arange(0, k, device="cuda:0", dtype=torch.int32) if k == n else torch.arange(1, k + 1, device="cuda:0", dtype=torch.int32) % n
    which = peers[torch.randint(0, peers.numel(), (N,), device="cuda:0", generator=g)]
    ids = (torch.randint(0, rows, (N,), device="cuda:0", dtype=torch.int32, generator=g) * n + which).to(torch.int32)
    for _ in range(2): table.gather(out, ids, N)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): table.gather(out, ids, N)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    ok = torch.equal(out[:64].cpu(), torch.stack([shards[int(v) % n][int(v) // n].cpu() for v in ids[:64].cpu()]))
    print(f"peers touched {peers.tolist()}: {ms:.3f} ms -> {N * F * 4 / ms / 1e6:.0f} GB/s  ok={ok}")
