"""Edge list -> global CSC on the device (csrc/ingest.cu) at the Reddit-shaped size (114M edges) and the products-shaped size,
against the host restatement (numpy stable argsort) on a 10M-edge sample. python tools/ingest_bench.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
nts = ge.load_package()
cs = nts.Cuda_Stream.on_torch_stream(0)
g = torch.Generator(device="cuda").manual_seed(1)
for name, V, E in (("reddit-shaped", 232965, 113_831_041), ("products-shaped", 2449029, 61_859_140)):
    pairs = torch.empty((E, 2), dtype=torch.int32, device="cuda")
    pairs[:, 0] = (torch.rand(E, generator=g, device="cuda").pow(1.6) * V).to(torch.int32).clamp_(0, V - 1)
    pairs[:, 1] = (torch.rand(E, generator=g, device="cuda").pow(2.0) * V).to(torch.int32).clamp_(0, V - 1)
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        gr = nts.FullyRepGraph(cs, V, edge_pairs=pairs)
        torch.cuda.synchronize(); dt = time.time() - t0
        del gr
    print(f"{name}: |V|={V} |E|={E}: device build from device-resident pairs {dt * 1e3:.1f} ms ({E / dt / 1e6:.0f} M edges/s, "
          f"{E * 8 / dt / 1e9:.0f} GB/s of edge list)")
    host = pairs.cpu().numpy().view(np.uint32)
    torch.cuda.synchronize(); t0 = time.time()
    gr = nts.FullyRepGraph(cs, V, edge_pairs=host)
    torch.cuda.synchronize(); dt = time.time() - t0
    print(f"{name}: the same from pageable host memory (H2D of {E * 8 / 1e9:.2f} GB included) {dt * 1e3:.1f} ms")
    del gr
    sample = host[:10_000_000]
    t0 = time.time(); nts.FullyRepGraph.build_csc_host(sample, V); dt = time.time() - t0
    print(f"{name}: host restatement (numpy stable argsort) on a 10M-edge sample {dt * 1e3:.0f} ms ({10 / dt:.1f} M edges/s)")
    del pairs, host
