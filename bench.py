#!/usr/bin/env python
"""bench.py -- sampled-edges/s of the sample-based hot path on a Reddit-shaped synthetic workload
(BASELINE.json configs[1]: 232,965 vertices, ~114.6M edges, F=602-128-41, fanout 25-10, batch 1024).

A step is one mini-batch through the hot path, exactly what toolkits/GCN_SAMPLE_*.hpp run per batch
between the dense layers:
    sample both layers (+ reindex, CSC, CSR, weights)         FastSampler::sample_gpu_fast
    gather X0 = features[source of the bottom layer]  (F=602)  load_feature_gpu
    aggregate forward bottom hop (F=602) and top hop (F=128)   SingleGPUAllSampleGraphOp::forward
    aggregate backward top hop (F=128)                         ::backward  (the bottom hop's backward into
                                                               X0 never runs: core/ntsContext.hpp:443)
    N > 1: one bucketed NCCL sum-allreduce of the dense weight gradients (602x128 + 128x41 floats)
The dense layers themselves are libtorch and outside the path; the top hop aggregates a resident
synthetic [S_0,128] activation instead.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo, one JSON line
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's own OpenMP CPU path
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

V, E_TARGET, F0, F1, NCLS = 232965, 114615892, 602, 128, 41
FANOUT, BATCH = [25, 10], 1024
TRAIN_FRAC = 0.66
SEED_GRAPH, SEED_SHUFFLE, SEED_SAMPLER = 0x5EED0001, 0x5EED0003, 0x5EED0004
REF_DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")


def reddit_shaped_graph(scale=1.0):
    """In-edge CSC with a power-law in-degree (mean ~492, min 1) and popularity-skewed sources."""
    v = max(1000, int(V * scale))
    e = int(E_TARGET * scale * scale) if scale != 1.0 else E_TARGET
    rng = np.random.default_rng(SEED_GRAPH)
    w = np.maximum(rng.random(v), 1e-6) ** -0.65
    deg = np.clip(np.floor(w / w.sum() * e), 1, v - 1).astype(np.int64)
    col_off = np.zeros(v + 1, np.int64)
    np.cumsum(deg, out=col_off[1:])
    total = int(col_off[-1])
    src = np.empty(total, np.uint32)
    chunk = 1 << 24
    for a in range(0, total, chunk):
        b = min(total, a + chunk)
        src[a:b] = np.minimum((rng.random(b - a) ** 1.6 * v).astype(np.int64), v - 1)
    return v, col_off.astype(np.uint32), src


def train_seeds(v):
    rng = np.random.default_rng(SEED_SHUFFLE)
    ids = rng.permutation(v)[: int(v * TRAIN_FRAC)].astype(np.uint32)
    return ids


def shard_seeds(ids, rank, world):
    """contiguous split of the training ids over the GPUs (toolkits/GAT_SAMPLE_ALL_MULTI.hpp:513-527)"""
    per = ids.size // world
    return ids[rank * per:(rank + 1) * per if rank < world - 1 else ids.size]


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region (NVML in-process, ~2 ms period)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.mask, self.stop_flag, self.max_sm = index, [], 0, False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.mask |= nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                pass
            time.sleep(0.002)

    def summary(self):
        self.stop_flag = True
        reasons = []
        if self.nv is not None:
            nv = self.nv
            for name, bit in (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                              ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap)):
                if self.mask & bit:
                    reasons.append(name)
        return {"sm_mhz": int(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm, "reasons": reasons,
                "samples": len(self.sm)}


def write_reference_inputs(td, v, col_off, src, seeds):
    """EDGE_FILE of (src,dst) u32 pairs in CSC order -> the reference rebuilds the identical CSC."""
    e = src.size
    pairs = np.empty((e, 2), np.uint32)
    pairs[:, 0] = src
    pairs[:, 1] = np.repeat(np.arange(v, dtype=np.uint32), np.diff(col_off.astype(np.int64)))
    ef, sf = os.path.join(td, "g.edge"), os.path.join(td, "seeds.u32")
    pairs.tofile(ef)
    seeds.tofile(sf)
    return ef, sf


def run_reference_driver(v, col_off, src, seeds, batches, warmup, threads):
    with tempfile.TemporaryDirectory() as td:
        ef, sf = write_reference_inputs(td, v, col_off, src, seeds)
        env = dict(os.environ, NTS_ORACLE_CPUS=str(threads + 1), OMP_NUM_THREADS=str(threads))
        out = subprocess.run([REF_DRIVER, "bench", ef, str(v), sf, str(BATCH), ",".join(map(str, FANOUT)), str(F0), str(F1),
                              str(batches), str(warmup)], capture_output=True, text=True, env=env, check=True).stdout
    return json.loads([l for l in out.splitlines() if l.startswith("{")][-1])


def cpu_metric(r):
    t = r["sample_s"] + r["gather_s"] + r["fwd_s"] + r["bwd_s"]
    return r["edges"] / t, t / max(r["batches"], 1) * 1e3


def config_dict(v, e, extra=None):
    c = {"workload": f"Reddit-shaped synthetic graph ({v} vertices, {e} edges, power-law in-degree), GCN_SAMPLE hot path: "
                     f"sample fanout 25-10 + reindex/CSC/CSR/weights + gather F=602 + aggregate fwd 602/128 + bwd 128",
         "batch": BATCH, "fanout": "25-10", "layers": "602-128-41",
         "l2": "inputs larger than L2: each batch gathers ~130K random rows (~313 MB) of a 561 MB HBM-resident table"}
    if extra:
        c.update(extra)
    return c


def main_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = os.cpu_count() or 1
    v, col_off, src = reddit_shaped_graph(args.scale)
    seeds = train_seeds(v)
    if not os.path.exists(REF_DRIVER):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_driver was not built"}))
        return
    r = run_reference_driver(v, col_off, src, seeds, args.steps, args.warmup, threads)
    val, ms = cpu_metric(r)
    cpu = {"value": val, "unit": "edges/s", "cores": threads, "kind": "reference",
           "sample": f"{r['batches']} mini-batches of {BATCH} seeds after {args.warmup} warm-up, all host threads (OpenMP); "
                     f"per-stage seconds sample/gather/fwd/bwd = {r['sample_s']:.3f}/{r['gather_s']:.3f}/{r['fwd_s']:.3f}/{r['bwd_s']:.3f}"}
    print(json.dumps({"impl": "reference", "metric": "sampled_edges_per_s", "value": val, "unit": "edges/s", "n_gpus": args.gpus,
                      "steps": r["batches"], "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": config_dict(v, int(src.size)), "cpu_baseline": cpu,
                      "e2e": {"value": val, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main_b200(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    nts = ge.load_package()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    v, col_off, src = reddit_shaped_graph(args.scale)
    e_total = int(src.size)
    all_seeds = train_seeds(v)
    my_seeds = shard_seeds(all_seeds, rank, world)
    n_steps = args.warmup + args.steps
    reps = -(-n_steps * BATCH // my_seeds.size)
    my_seeds = np.tile(my_seeds, reps)[: n_steps * BATCH]

    stream = torch.cuda.Stream(dev)
    with torch.cuda.stream(stream):
        cs = nts.Cuda_Stream(local, stream)
        graph = nts.FullyRepGraph(cs, v, column_offset=col_off, row_indices=src)
        sampler = nts.FastSampler(graph, my_seeds, 2, BATCH, FANOUT, cuda_stream=cs, build_csr=True, rng_seed=SEED_SAMPLER + rank)
        lib, check, ptr = nts._capi.lib(), nts._capi.check, nts._capi.ptr
        gen = torch.Generator(device=dev).manual_seed(0x5EED0002)
        table = torch.rand((v, F0), generator=gen, device=dev) * 2 - 1          # HBM-resident feature table
        cap_s1 = min(BATCH * 25 * 10, v)
        cap_s0 = min(BATCH * 25, v)
        x0 = torch.empty((cap_s1, F0), device=dev)
        y1 = torch.empty((cap_s0, F0), device=dev)
        h1 = torch.rand((cap_s0, F1), generator=gen, device=dev)               # stands in for relu(Y1 W1)
        y0 = torch.empty((BATCH, F1), device=dev)
        dy0 = torch.rand((BATCH, F1), generator=gen, device=dev)
        dh1 = torch.empty((cap_s0, F1), device=dev)
        grads = torch.zeros(F0 * F1 + F1 * NCLS, device=dev)                    # dense W gradients, one bucket
        seeds_dev = torch.from_numpy(my_seeds.view(np.int32)).to(dev)
        seeds_pin = torch.from_numpy(my_seeds.view(np.int32)).pin_memory()
        y0_host = torch.empty((BATCH, F1)).pin_memory()
        C = nts._capi.C
        views = (nts._capi.LayerView * 2)()
        # one synchronous batch: fixes the arena pointers of both layers (they never change afterwards)
        check(lib.nb_sampler_sample(sampler._samplers[0], ptr(seeds_pin[:BATCH]), BATCH, 0, SEED_SAMPLER + rank, 0,
                                    nts.WeightType.Sum, None, 0xFFFFFFFF, views, 1))
        top, bot = views[0], views[1]
        nd, ne, ns = [[C.c_void_p() for _ in range(2)] for _ in range(3)]
        caps = [[C.c_uint32() for _ in range(3)] for _ in range(2)]
        for l in range(2):
            check(lib.nb_sampler_sizes_dev(sampler._samplers[0], l, C.byref(nd[l]), C.byref(ne[l]), C.byref(ns[l]),
                                           C.byref(caps[l][0]), C.byref(caps[l][1]), C.byref(caps[l][2])))
        assert caps[1][2].value <= cap_s1 and caps[0][2].value <= cap_s0
        sizes_pin = torch.zeros((n_steps, 8), dtype=torch.int32).pin_memory()   # LayerMeta of the bottom layer per step
        sizes_top = torch.zeros((n_steps, 8), dtype=torch.int32).pin_memory()
        fast = nts.FastSampler(graph, my_seeds, 2, BATCH, FANOUT, cuda_stream=cs, build_csr=True, rng_seed=SEED_SAMPLER + rank)
        op_bot_cls = nts.SingleGPUAllSampleGraphOp

        ev = lambda: torch.cuda.Event(enable_timing=True)
        kern_ev = {"gather": [], "agg_fwd_602": []}

        def step_async(i, timed):
            """value: inputs resident in HBM, no host synchronisation anywhere in the step (sizes stay on the device)."""
            check(lib.nb_sampler_sample(sampler._samplers[0], ptr(seeds_dev[i * BATCH:(i + 1) * BATCH]), BATCH, 1,
                                        SEED_SAMPLER + rank, i, nts.WeightType.Sum, None, 0xFFFFFFFF, None, 0))
            if timed:
                a, b, c = ev(), ev(), ev()
                a.record(stream)
            check(lib.nb_gather_rows_dyn(cs._h, ptr(x0), ptr(table), bot.source, ns[1], caps[1][2], F0, F0, F0))
            if timed:
                b.record(stream)
            check(lib.nb_aggregate_csc_fwd_dyn(cs._h, ptr(x0), ptr(y1), bot.edge_weight_forward, bot.row_indices,
                                               bot.column_offset, nd[1], caps[1][0], F0, F0, F0))
            if timed:
                c.record(stream)
                kern_ev["gather"].append((a, b))
                kern_ev["agg_fwd_602"].append((b, c))
            check(lib.nb_aggregate_csc_fwd_dyn(cs._h, ptr(h1), ptr(y0), top.edge_weight_forward, top.row_indices,
                                               top.column_offset, nd[0], caps[0][0], F1, F1, F1))
            check(lib.nb_aggregate_csr_bwd_dyn(cs._h, ptr(dy0), ptr(dh1), top.edge_weight_backward, top.row_offset,
                                               top.column_indices, ns[0], caps[0][2], F1, F1, F1))
            check(lib.nb_memcpy_d2h(cs._h, ptr(sizes_pin[i]), nd[1].value, 32, 0))
            check(lib.nb_memcpy_d2h(cs._h, ptr(sizes_top[i]), nd[0].value, 32, 0))
            if world > 1:
                dist.all_reduce(grads)

        def step_api(i, timed):
            """e2e: the call sequence a user of the reference-shaped API makes (host seeds in, sizes and the batch's
            top-layer output back on the host), including its synchronisation on the sampled sizes."""
            fast.work_offset = i * BATCH
            sg = fast.sample_gpu_fast(BATCH)                        # H2D seeds; syncs for the sizes
            t, bt = sg.sampled_sgs
            xx = x0[:bt.src_size]
            fast.load_feature_gpu(cs, sg, xx, table)
            yy1 = op_bot_cls(sg, 1, cs).forward(xx)
            op_top = op_bot_cls(sg, 0, cs)
            yy0 = op_top.forward(h1[:t.src_size])
            op_top.backward(dy0)
            if world > 1:
                dist.all_reduce(grads)
            y0_host.copy_(yy0, non_blocking=True)
            stream.synchronize()
            sizes_pin[i, 0], sizes_pin[i, 1], sizes_pin[i, 2] = bt.v_size, bt.e_size, bt.src_size
            sizes_top[i, 1] = t.e_size
            del yy1

        clock_box = [None]

        def run(from_host, sample_clocks=False):
            step = step_api if from_host else step_async
            for k in kern_ev.values():
                k.clear()
            for i in range(args.warmup):
                step(i, False)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            launches0 = cs.launch_count() + sum(c.launch_count() for c in set(fast.cs_array) if c is not cs)
            clocks = ClockSampler(local) if sample_clocks else None
            if clocks:
                clocks.start()
            t0, t1 = ev(), ev()
            t0.record(stream)
            for i in range(args.warmup, n_steps):
                step(i, True)
            t1.record(stream)
            torch.cuda.synchronize()
            if clocks:
                clock_box[0] = clocks.summary()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            ms = t0.elapsed_time(t1)
            sp, st_ = sizes_pin[args.warmup:n_steps].numpy().astype(np.int64), sizes_top[args.warmup:n_steps].numpy().astype(np.int64)
            work = {"edges": int(sp[:, 1].sum() + st_[:, 1].sum()), "V1": int(sp[:, 0].sum()), "E1": int(sp[:, 1].sum()),
                    "S1": int(sp[:, 2].sum())}
            return ms, cs.launch_count() - launches0, work, {k: sum(a.elapsed_time(b) for a, b in v) / max(len(v), 1)
                                                            for k, v in kern_ev.items()}

        ms, launches, work, kms = run(False, sample_clocks=True)
        clk = clock_box[0]
        ms_e2e, _, work_e2e, _ = run(True)

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    ms_max, ms_e2e_max = reduce_max(ms), reduce_max(ms_e2e)
    edges_all, edges_e2e_all = reduce_sum(work["edges"]), reduce_sum(work_e2e["edges"])
    launches_all = reduce_sum(launches)
    value = edges_all / (ms_max * 1e-3)
    e2e_value = edges_e2e_all / (ms_e2e_max * 1e-3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        n = args.steps
        S1, E1, V1 = work["S1"] / n, work["E1"] / n, work["V1"] / n
        bytes_gather = S1 * (4 + 8 * F0)
        bytes_agg = E1 * (8 + 4 * F0) + 4 * (V1 + 1) + 4 * V1 * F0
        kernels = {"gather_rows(F=602)": {"ms": kms["gather"], "algorithmic_bytes": bytes_gather,
                                          "gbs": bytes_gather / (kms["gather"] * 1e-3) / 1e9},
                   "segment_reduce_fwd(F=602)": {"ms": kms["agg_fwd_602"], "algorithmic_bytes": bytes_agg,
                                                 "gbs": bytes_agg / (kms["agg_fwd_602"] * 1e-3) / 1e9}}
        dom = max(kernels, key=lambda k: kernels[k]["ms"])
        roof = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["gbs"] / peak, "traffic": None, "peak_source": peak_src,
                "kernels": {k: {"gbs": round(x["gbs"], 1), "frac": round(x["gbs"] / peak, 3), "ms": round(x["ms"], 4)} for k, x in kernels.items()}}
        traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(traffic_file):
            try:
                roof["traffic"] = json.load(open(traffic_file)).get(dom)
            except Exception:
                pass
        cpu = None
        if world == 1 and not args.no_cpu_baseline and os.path.exists(REF_DRIVER):
            threads = os.cpu_count() or 1
            try:
                r = run_reference_driver(v, col_off, src, all_seeds, args.cpu_batches, 2, threads)
                val, _ = cpu_metric(r)
                cpu = {"value": val, "unit": "edges/s", "cores": threads, "kind": "reference",
                       "sample": f"{r['batches']} mini-batches of {BATCH} seeds of the same workload through oracle/_ref/ref_driver "
                                 f"(the reference's own OpenMP sample_fast/get_feature/MiniBatchFuseOp), all host threads; "
                                 f"sample/gather/fwd/bwd s = {r['sample_s']:.3f}/{r['gather_s']:.3f}/{r['fwd_s']:.3f}/{r['bwd_s']:.3f}"}
            except Exception as ex:  # the checker must never take the bench down
                cpu = {"value": None, "unit": "edges/s", "cores": threads, "kind": "reference", "sample": f"failed: {ex}"}
        line = {"metric": "sampled_edges_per_s", "value": value, "unit": "edges/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_dict(v, e_total, {"per_gpu_batch": BATCH, "parallelism": f"dp{world}: seeds sharded, "
                                                   "one bucketed NCCL allreduce of dense grads per step" if world > 1 else "single GPU",
                                                   "avg_E_per_step": work["edges"] / n, "avg_S1": S1, "avg_E1": E1, "avg_V1": V1,
                                                   "epoch_ms_est": (ms_max / args.steps) * (all_seeds.size / BATCH / world)}),
                "e2e": {"value": e2e_value, "unit": "edges/s", "h2d_bytes_per_step": BATCH * 4,
                        "d2h_bytes_per_step": BATCH * F1 * 4 + 2 * 32, "ms_per_step": ms_e2e_max / args.steps},
                "gpu_launches": int(launches_all), "clocks": clk, "roofline": roof, "cpu_baseline": cpu}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the graph (debug only; the metric is quoted at 1.0)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-batches", type=int, default=20)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        main_reference(args)
    else:
        main_b200(args)
