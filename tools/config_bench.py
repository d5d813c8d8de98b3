"""Per-config measurements for BASELINE.json configs[2] (products-shaped, hotness-aware cache, 1 GPU) and configs[4]
(papers100M-shaped, feature table / hot cache sharded over NVLink, host-streamed cold rows, N GPUs). These are not bench.py
lines (the headline is configs[1]); they document where the time goes on the other shapes.

    python tools/config_bench.py --config products [--scale 1.0]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/config_bench.py --config papers --scale 0.125

Each step = sample (fanout 25-10, batch 1024) -> gather X0 by the listed variant -> aggregate fwd (bottom F, top 128) -> bwd (top).
Times are CUDA-event times on the launching stream, max over ranks; one JSON line per variant."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402


def power_law_graph(V, E, seed):
    """in-edge CSC generated on the device: power-law in-degree, skewed sources (the recipe of tests/test_gpu_parity.py)"""
    g = torch.Generator(device="cuda").manual_seed(seed)
    w = torch.rand(V, generator=g, device="cuda").clamp_min(1e-6).pow(-0.65)
    deg = (w / w.sum() * E).floor().clamp_(1, V - 1).to(torch.int64)
    co = torch.zeros(V + 1, dtype=torch.int64, device="cuda")
    co[1:] = deg.cumsum(0)
    total = int(co[-1])
    src = torch.empty(total, dtype=torch.int32, device="cuda")
    chunk = 1 << 27
    for a in range(0, total, chunk):
        b = min(total, a + chunk)
        src[a:b] = (torch.rand(b - a, generator=g, device="cuda").pow(1.6) * V).to(torch.int64).clamp_(0, V - 1).to(torch.int32)
    out_deg = torch.bincount(src.long(), minlength=V)
    co_np, src_np = co.to(torch.int32).cpu().numpy().view(np.uint32), src.cpu().numpy().view(np.uint32)
    del src, co, w, deg
    torch.cuda.empty_cache()
    return co_np, src_np, out_deg


class Runner:
    def __init__(self, nts, cs, stream, graph, seeds, F, hidden, world):
        self.nts, self.cs, self.stream, self.F, self.hidden, self.world = nts, cs, stream, F, hidden, world
        self.batch = 1024
        self.sampler = nts.FastSampler(graph, seeds, 2, self.batch, [25, 10], cuda_stream=cs, bottom_csr=False)
        cap = self.batch * 25 * 10 + self.batch * 25 + self.batch
        self.x0 = torch.empty((cap, F), device="cuda")
        self.y1 = torch.empty((self.batch * 26, F), device="cuda")
        self.h1 = torch.randn((self.batch * 26, hidden), device="cuda")
        self.y0 = torch.empty((self.batch, hidden), device="cuda")
        self.dh1 = torch.empty((self.batch * 26, hidden), device="cuda")
        self.dy0 = torch.randn((self.batch, hidden), device="cuda")

    def next_batch(self, cache_flag=None):
        if not self.sampler.sample_not_finished():
            self.sampler.restart()
        return self.sampler.sample_gpu_fast(self.batch, 0, self.nts.WeightType.Sum, cache_flag, 0 if cache_flag is not None else 0xFFFFFFFF)

    def aggregate(self, sg, x0):
        cs, (top, bot) = self.cs, sg.sampled_sgs
        cs.Gather_By_Dst_From_Src_Spmm(x0, self.y1, bot.dev_edge_weight_forward, bot.dev_row_indices, bot.dev_column_offset,
                                       bot.src_size, edges=bot.e_size, batch_size=bot.v_size, feature_size=self.F, with_weight=True)
        cs.Gather_By_Dst_From_Src_Spmm(self.h1, self.y0, top.dev_edge_weight_forward, top.dev_row_indices, top.dev_column_offset,
                                       top.src_size, edges=top.e_size, batch_size=top.v_size, feature_size=self.hidden, with_weight=True)
        cs.Gather_By_Src_From_Dst_Spmm(self.dy0, self.dh1, top.dev_edge_weight_backward, top.dev_row_offset, top.dev_column_indices,
                                       top.v_size, edges=top.e_size, batch_size=top.src_size, feature_size=self.hidden, with_weight=True)

    def run(self, name, gather_fn, steps, warmup, cache_flag=None, extra=None):
        """gather_fn(sg, x0) -> number of cold rows (or None). Returns the JSON record."""
        ev = lambda: torch.cuda.Event(enable_timing=True)
        tot_e = tot_s = tot_cold = 0
        g_ms = s_ms = 0.0
        for it in range(warmup + steps):
            if it == warmup:
                torch.cuda.synchronize()
                if self.world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                t0 = ev(); t0.record(self.stream)
            a, b, c = ev(), ev(), ev()
            a.record(self.stream)
            sg = self.next_batch(cache_flag)
            b.record(self.stream)
            bot = sg.sampled_sgs[1]
            cold = gather_fn(sg, self.x0[:bot.src_size])
            c.record(self.stream)
            self.aggregate(sg, self.x0[:bot.src_size])
            if it >= warmup:
                tot_e += sum(l.e_size for l in sg.sampled_sgs)
                tot_s += bot.src_size
                tot_cold += cold or 0
                c.synchronize()
                s_ms += a.elapsed_time(b)
                g_ms += b.elapsed_time(c)
        t1 = ev(); t1.record(self.stream)
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1)
        stats = torch.tensor([ms, g_ms, s_ms], dtype=torch.float64, device="cuda")
        sums = torch.tensor([tot_e, tot_s, tot_cold], dtype=torch.float64, device="cuda")
        if self.world > 1:
            dist.all_reduce(stats, op=dist.ReduceOp.MAX)
            dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        ms, g_ms, s_ms = stats.tolist()
        tot_e, tot_s, tot_cold = sums.tolist()
        rec = {"variant": name, "n_gpus": self.world, "steps": steps, "ms_per_step": ms / steps, "sampled_edges_per_s": tot_e / (ms / 1e3),
               "sample_ms": s_ms / steps, "gather_ms": g_ms / steps, "rows_per_step_per_gpu": tot_s / steps / self.world,
               "gather_GBps_per_gpu_algorithmic": (tot_s / self.world) * (4 + 8 * self.F) / (g_ms / 1e3) / 1e9,
               "cold_fraction": tot_cold / max(tot_s, 1.0)}
        rec.update(extra or {})
        return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", choices=["products", "papers"], required=True)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--cache-rate", type=float, default=0.1, help="FEATURE_CACHE_RATE: fraction of the vertices kept hot (top out-degree)")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1 or args.config == "papers":
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29733")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    nts = ge.load_package()
    from sample_based_gnn_b200 import dist as nd
    stream = torch.cuda.Stream()
    records = []
    with torch.cuda.stream(stream):
        cs = nts.Cuda_Stream(local, stream)
        if args.config == "products":
            V, E, F = int(2449029 * args.scale), int(61859140 * args.scale), 100
        else:
            V, E, F = int(111059956 * args.scale), int(1615685872 * args.scale), 128
        co, ri, out_deg = power_law_graph(V, E, 0xBEEF if args.config == "products" else 0xFACE)   # same graph on every rank
        graph = nts.FullyRepGraph(cs, V, column_offset=co, row_indices=ri)
        del co, ri
        n_train = max(4096, int(V * (0.08 if args.config == "products" else 0.011)))
        train = np.random.default_rng(3).permutation(V)[:n_train].astype(np.uint32)
        seeds = nd.shard_seeds(train, rank, world) if world > 1 else train
        R = Runner(nts, cs, stream, graph, seeds, F, 128, world)
        shape = {"config": args.config, "V": V, "E": int(E), "F": F, "scale": args.scale, "batch": 1024, "fanout": "25-10"}
        # hot set: top cache_rate*V by out-degree (GS_SAMPLE_CACHE.hpp cache_high_degree / GS_SAMPLE_PC_MULTI.hpp:916-1015)
        n_hot = int(V * args.cache_rate)
        hot = torch.sort(torch.topk(out_deg, n_hot).indices).values
        hashmap = torch.full((V,), -1, dtype=torch.int32, device="cuda")
        hashmap[hot] = torch.arange(n_hot, dtype=torch.int32, device="cuda")
        del out_deg
        g = torch.Generator(device="cuda").manual_seed(11)

        def chunks(n, step=1 << 22):
            for a in range(0, n, step):
                yield a, min(n, a + step)

        if args.config == "products":
            table = torch.rand((V, F), device="cuda", generator=g)                      # whole table resident in HBM (0.98 GB)
            cache_table = table[hot].contiguous()
            host_pinned = torch.empty((V, F), dtype=torch.float32, pin_memory=True)
            host_pinned.copy_(table)
            torch.cuda.synchronize()
            hits = torch.zeros(1, dtype=torch.int32, device="cuda")
            def plain(sg, x):
                R.sampler.load_feature_gpu(cs, sg, x, table)

            def cached(sg, x):
                R.sampler.load_feature_gpu_cache(cs, sg, x, table, cache_table, hashmap, hits)

            def zero_copy(sg, x):      # the reference's layout: the kernel reads the mapped pinned host table over PCIe
                R.sampler.load_feature_gpu(cs, sg, x, host_pinned)

            def cached_zero_copy(sg, x):
                R.sampler.load_feature_gpu_cache(cs, sg, x, host_pinned, cache_table, hashmap, hits)
            records.append(R.run("hbm_table_plain_gather", plain, args.steps, args.warmup, extra=shape))
            records.append(R.run("hot_cache+cold_from_hbm_table", cached, args.steps, args.warmup, extra=shape))
            records.append(R.run("zero_copy_host_table(reference layout)", zero_copy, max(10, args.steps // 4), 2, extra=shape))
            records.append(R.run("hot_cache+cold_zero_copy_host(reference layout)", cached_zero_copy, max(10, args.steps // 4), 2,
                                 extra=dict(shape, cache_rate=args.cache_rate)))
            stage = nts.ColdStage(cs, host_pinned, max_rows=R.x0.shape[0])

            def staged(sg, x):
                bot = sg.sampled_sgs[1]
                stage.submit(0, bot.dev_source, bot.src_size, hashmap)
                return stage.gather(0, x, cache_table, hashmap, bot.dev_source)
            records.append(R.run("hot_cache+cold_staged_from_host(serial)", staged, args.steps, args.warmup, extra=shape))
            del stage
        else:
            # variant A: the whole table lives in HBM, row-sharded over the ranks, read over NVLink inside the gather kernel
            n_local = (V - rank + world - 1) // world
            mine = torch.empty((n_local, F), device="cuda")
            for a, b in chunks(n_local):
                mine[a:b] = torch.rand((b - a, F), device="cuda", generator=g)
            full = nd.ShardedTable(cs, mine, V, F)
            del mine
            torch.cuda.empty_cache()

            def sharded(sg, x):
                bot = sg.sampled_sgs[1]
                full.gather(x, bot.dev_source, bot.src_size)
            records.append(R.run("whole_table_sharded_in_hbm", sharded, args.steps, args.warmup,
                                 extra=dict(shape, hbm_table_GB_per_gpu=n_local * F * 4 / 1e9)))
            # variant B: hot cache (top out-degree) sharded over the ranks + cold rows staged from this rank's host table
            hot_rows = torch.empty((len(range(rank, n_hot, world)), F), device="cuda")
            hot_rows.uniform_(generator=g)
            hot_table = nd.ShardedTable(cs, hot_rows, n_hot, F)
            host_table = torch.empty((V, F), dtype=torch.float32)           # pageable host memory, this rank's copy
            buf = torch.empty((1 << 22, F), device="cuda")
            for a, b in chunks(V):
                buf[:b - a].uniform_(generator=g)
                host_table[a:b].copy_(buf[:b - a])
            del buf
            stage = nts.ColdStage(cs, host_table, max_rows=R.x0.shape[0])

            def tiered(sg, x):
                bot = sg.sampled_sgs[1]
                stage.submit(0, bot.dev_source, bot.src_size, hashmap)
                return stage.gather_table(0, x, hot_table.table, hashmap, bot.dev_source)
            records.append(R.run("hot_cache_sharded+cold_staged_from_host(serial)", tiered, max(10, args.steps // 4), 2,
                                 extra=dict(shape, cache_rate=args.cache_rate, hot_cache_GB_per_gpu=hot_rows.numel() * 4 / 1e9)))

            # the same, software pipelined: batch i+1 is sampled and its cold rows are staged while batch i is merged + aggregated
            def pipelined(steps, warmup):
                ev = lambda: torch.cuda.Event(enable_timing=True)
                sampler2 = nts.FastSampler(graph, seeds, 2, 1024, [25, 10], pipeline_num=2, cuda_stream=cs, bottom_csr=False)
                tot_e = tot_s = tot_cold = 0

                def issue(slot):
                    if not sampler2.sample_not_finished():
                        sampler2.restart()
                    sg = sampler2.sample_gpu_fast(1024, slot)
                    bot = sg.sampled_sgs[1]
                    stage.submit(slot, bot.dev_source, bot.src_size, hashmap)
                    return sg
                cur = issue(0)
                for it in range(warmup + steps):
                    if it == warmup:
                        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
                        t0 = ev(); t0.record(stream)
                    nxt = issue((it + 1) % 2)
                    bot = cur.sampled_sgs[1]
                    cold = stage.gather_table(it % 2, R.x0[:bot.src_size], hot_table.table, hashmap, bot.dev_source)
                    R.aggregate(cur, R.x0[:bot.src_size])
                    if it >= warmup:
                        tot_e += sum(l.e_size for l in cur.sampled_sgs); tot_s += bot.src_size; tot_cold += cold
                    cur = nxt
                t1 = ev(); t1.record(stream); torch.cuda.synchronize()
                stage.gather_table((warmup + steps) % 2, R.x0[:cur.sampled_sgs[1].src_size], hot_table.table, hashmap, cur.sampled_sgs[1].dev_source)
                torch.cuda.synchronize()
                stats = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device="cuda")
                sums = torch.tensor([tot_e, tot_s, tot_cold], dtype=torch.float64, device="cuda")
                dist.all_reduce(stats, op=dist.ReduceOp.MAX); dist.all_reduce(sums, op=dist.ReduceOp.SUM)
                ms = float(stats[0]); e, s, c = sums.tolist()
                return dict(shape, variant="hot_cache_sharded+cold_staged_from_host(2-slot pipeline)", n_gpus=world, steps=steps,
                            ms_per_step=ms / steps, sampled_edges_per_s=e / (ms / 1e3), rows_per_step_per_gpu=s / steps / world,
                            cold_fraction=c / max(s, 1.0), cold_GBps_per_gpu=(c / world) * F * 4 / (ms / 1e3) / 1e9,
                            cache_rate=args.cache_rate)
            records.append(pipelined(max(10, args.steps // 4), 2))
            del stage
            torch.cuda.synchronize()
            hot_table.close(); full.close()
    if rank == 0:
        for r in records:
            print("CONFIG_BENCH " + json.dumps(r))
        if args.out:
            with open(args.out, "a") as f:
                for r in records:
                    f.write(json.dumps(r) + "\n")
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
