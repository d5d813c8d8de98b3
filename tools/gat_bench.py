"""Fused GAT layer (nb_gat_fwd / nb_gat_bwd) vs the legacy five-op chain on the Reddit-shaped workload (BASELINE.json configs[3]
shape: hidden 128, one head, fanout 25-10, batch 1024). python tools/gat_bench.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402
import bench as B  # noqa: E402

nts = ge.load_package()
for kv in sys.argv[1:]:       # name=value pairs for nb_set_option (A/B of tuning knobs)
    name, value = kv.split("=")
    nts._capi.check(nts._capi.lib().nb_set_option(name.encode(), int(value)))
    print("option", name, value)
v, col_off, src = B.reddit_shaped_graph(1.0)
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    cs = nts.Cuda_Stream(0, stream)
    graph = nts.FullyRepGraph(cs, v, column_offset=col_off, row_indices=src)
    fs = nts.FastSampler(graph, B.train_seeds(v), 2, 1024, [25, 10], cuda_stream=cs, merge_src_dst=True, build_csr=True)
    sg = fs.sample_gpu_fast(1024, weightType=nts.WeightType.None_)
    F = 128

    def timeit(name, fn, nbytes, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        for i in range(reps):
            ev[i].record(stream); fn()
        ev[reps].record(stream); torch.cuda.synchronize()
        ms = float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]))
        print(f"{name:58s} {ms * 1e3:9.1f} us {nbytes / ms / 1e6:9.1f} GB/s(alg)")

    for hop in (1, 0):
        lay = sg.sampled_sgs[hop]
        S, E, V1 = lay.src_size, lay.e_size, lay.v_size
        print(f"hop {hop}: S={S} E={E} V={V1}")
        h = torch.randn((S, F), device="cuda"); att = torch.randn(2 * F, device="cuda") * 0.1
        dout = torch.randn((V1, F), device="cuda")
        op = nts.GATFusedOp(sg, hop, cs)
        fwd_bytes = E * (4 + 4 * F) + 4 * (V1 + 1) + 4 * (S + V1) + 4 * V1 * F + 4 * E + S * 4 * F   # + node-score pass over H
        timeit(f"fused fwd  hop {hop}", lambda: op.forward(h, att), fwd_bytes)
        timeit(f"fused bwd  hop {hop}", lambda: op.backward(h, att, dout), 2 * E * 4 * F + 2 * S * 4 * F + V1 * 4 * F)

        def legacy():
            msg = nts.BatchGPUSrcDstScatterOp(sg, hop, cs).forward(h)
            m = torch.nn.functional.leaky_relu(msg @ att.view(-1, 1), 0.2)
            a = nts.BatchGPUEdgeSoftMax(sg, hop, cs).forward(m)
            return nts.BatchGPUAggregateDst(sg, hop, cs).forward(msg[:, :F] * a)
        timeit(f"legacy 5-op chain fwd hop {hop} (reference composition)", legacy, fwd_bytes)
