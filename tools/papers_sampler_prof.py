"""Sampling time on the papers100M-shaped topology at full |V| (111M vertices, ~1.6B edges): the two-level dedup bitmap keeps a batch
O(|V|/1024 + S + E). Prints the mean time of one batch (CUDA events around the sampler's graph launch). python tools/papers_sampler_prof.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402
import bench as B  # noqa: E402

nts = ge.load_package()
V, E = 111059956, 1615685872
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    cs = nts.Cuda_Stream(0, stream)
    co, src = B.power_law_graph_gpu(torch, V, E, 0xFACE)
    graph = nts.FullyRepGraph(cs, V, column_offset=co, row_indices=src)
    del co, src
    torch.cuda.empty_cache()
    seeds = np.random.default_rng(3).permutation(V)[:64 * 1024].astype(np.uint32)
    for fused in (1, 0):
        nts._capi.check(nts._capi.lib().nb_set_option(b"sampler_fused", fused))
        fs = nts.FastSampler(graph, seeds, 2, 1024, [25, 10], cuda_stream=cs, bottom_csr=False)
        for _ in range(4):
            sg = fs.sample_gpu_fast(1024)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
        for i in range(20):
            ev[i].record(stream)
            fs.sample_gpu_fast(1024, sync=False)
        ev[20].record(stream)
        torch.cuda.synchronize()
        ms = float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(20)]))
        print(f"papers100M-shaped, |V|={V}: one 2-layer batch (E={[l.e_size for l in sg.sampled_sgs]}, S={[l.src_size for l in sg.sampled_sgs]}) "
              f"sampler_fused={fused}: {ms * 1e3:.1f} us")
        del fs
