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


def run_oracle_port(v, col_off, src, seeds, batches, warmup):
    """Fallback when oracle/_ref/ref_driver is not available: the scalar C restatement (oracle/oracle.c), 1 thread.
    Same stages as the reference driver's bench mode; returns the same dict."""
    import oracle
    ind = np.maximum(np.diff(col_off.astype(np.int64)), 1).astype(np.uint32)
    outd = np.maximum(np.bincount(src, minlength=v), 1).astype(np.uint32)
    table = np.ones((v, F0), np.float32)
    acc = dict(batches=0, threads=1, sample_s=0.0, gather_s=0.0, fwd_s=0.0, bwd_s=0.0, edges=0, rows=0)
    for b in range(batches + warmup):
        sd = seeds[b * BATCH:(b + 1) * BATCH]
        t0 = time.time()
        lay = oracle.sample_batch(sd, col_off, src, FANOUT, v, ind, outd, seed=b)
        t1 = time.time()
        x0 = oracle.gather_rows(table, lay[1]["source"])
        t2 = time.time()
        y1 = oracle.aggregate_fwd(x0, lay[1]["column_offset"], lay[1]["row_indices"], lay[1]["e_w_f"])
        h1 = np.ones((lay[0]["source"].size, F1), np.float32)
        y0 = oracle.aggregate_fwd(h1, lay[0]["column_offset"], lay[0]["row_indices"], lay[0]["e_w_f"])
        t3 = time.time()
        oracle.aggregate_bwd_csr(np.ones_like(y0), lay[0]["row_offset"], lay[0]["column_indices"], lay[0]["e_w_b"])
        t4 = time.time()
        del y1
        if b >= warmup:
            acc["batches"] += 1
            acc["sample_s"] += t1 - t0; acc["gather_s"] += t2 - t1; acc["fwd_s"] += t3 - t2; acc["bwd_s"] += t4 - t3
            acc["edges"] += int(lay[0]["sample_ans"].size + lay[1]["sample_ans"].size)
            acc["rows"] += int(lay[1]["source"].size)
    return acc


def cpu_baseline_run(v, col_off, src, seeds, batches, warmup):
    """(result dict, kind, cores): the reference's own OpenMP path when its driver is present, else the C port"""
    if os.path.exists(REF_DRIVER):
        threads = os.cpu_count() or 1
        return run_reference_driver(v, col_off, src, seeds, batches, warmup, threads), "reference", threads
    return run_oracle_port(v, col_off, src, seeds, min(batches, 5), min(warmup, 1)), "port", 1


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
    v, col_off, src = reddit_shaped_graph(args.scale)
    seeds = train_seeds(v)
    steps = min(args.steps, 100)   # bounded sample: the CPU path takes ~30-60 ms per mini-batch
    r, kind, threads = cpu_baseline_run(v, col_off, src, seeds, steps, args.warmup)
    val, ms = cpu_metric(r)
    cpu = {"value": val, "unit": "edges/s", "cores": threads, "kind": kind,
           "sample": f"{r['batches']} mini-batches of {BATCH} seeds after {args.warmup} warm-up, "
                     + ("oracle/_ref/ref_driver = the reference's own OpenMP sample_fast/get_feature/MiniBatchFuseOp on all host threads; "
                        if kind == "reference" else "oracle/oracle.c scalar port (reference driver not built); ")
                     + f"per-stage seconds sample/gather/fwd/bwd = {r['sample_s']:.3f}/{r['gather_s']:.3f}/{r['fwd_s']:.3f}/{r['bwd_s']:.3f}"}
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
    C = nts._capi.C
    lib, check, ptr = nts._capi.lib(), nts._capi.check, nts._capi.ptr

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ["NCCL_DEBUG"] = "WARN"   # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)
    v, col_off, src = reddit_shaped_graph(args.scale)
    e_total = int(src.size)
    all_seeds = train_seeds(v)
    my_seeds = shard_seeds(all_seeds, rank, world)
    n_steps = args.warmup + args.steps
    reps = -(-n_steps * BATCH // my_seeds.size)
    my_seeds = np.tile(my_seeds, reps)[: n_steps * BATCH]
    P = max(1, args.pipeline)
    PITCH = args.pitch if args.pitch else F0          # row pitch of the 602-wide tensors, in floats

    for kv in args.opt:
        name, value = kv.split("=")
        check(lib.nb_set_option(name.encode(), int(value)))
    st_sample, st_train = torch.cuda.Stream(dev, priority=args.sample_priority), torch.cuda.Stream(dev)
    cs_sample, cs_train = nts.Cuda_Stream(local, st_sample), nts.Cuda_Stream(local, st_train)
    # e2e path: the host waits for every batch's sampled sizes, so sampling is on its critical path and gets a high-priority
    # stream (its small kernels are scheduled ahead of the resident gather / aggregation blocks of the previous batch). In the
    # device-resident path nothing waits for the sampler, and a normal-priority stream leaves the aggregation undisturbed.
    st_sample_api = torch.cuda.Stream(dev, priority=-1)
    cs_sample_api = nts.Cuda_Stream(local, st_sample_api)
    with torch.cuda.stream(st_train):
        graph = nts.FullyRepGraph(cs_sample, v, column_offset=col_off, row_indices=src)
        # one sampler (arena) per pipeline slot, all on the sampling stream (the reference's PIPELINE_NUM SampledSubgraphs)
        # bottom_csr=False: the bottom hop's backward never runs in the GCN toolkits (core/ntsContext.hpp:443), so its CSR is not built
        sampler = nts.FastSampler(graph, my_seeds, 2, BATCH, FANOUT, pipeline_num=P, cuda_stream=[cs_sample] * P, build_csr=True,
                                  bottom_csr=False, rng_seed=SEED_SAMPLER + rank)
        fast = nts.FastSampler(graph, my_seeds, 2, BATCH, FANOUT, pipeline_num=2, cuda_stream=[cs_sample_api] * 2, build_csr=True,
                               bottom_csr=False, rng_seed=SEED_SAMPLER + rank)
        api_ev = [dict(sampled=torch.cuda.Event(), consumed=torch.cuda.Event()) for _ in range(2)]
        gen = torch.Generator(device=dev).manual_seed(0x5EED0002)
        table = torch.zeros((v, PITCH), device=dev)                               # HBM-resident feature table
        table[:, :F0] = torch.rand((v, F0), generator=gen, device=dev) * 2 - 1
        cap_s1, cap_s0 = min(BATCH * 25 * 10, v), min(BATCH * 25, v)
        x0 = torch.zeros((cap_s1, PITCH), device=dev)
        y1 = torch.zeros((cap_s0, PITCH), device=dev)
        h1 = torch.rand((cap_s0, F1), generator=gen, device=dev)                 # stands in for relu(Y1 W1)
        y0 = torch.empty((BATCH, F1), device=dev)
        dy0 = torch.rand((BATCH, F1), generator=gen, device=dev)
        dh1 = torch.empty((cap_s0, F1), device=dev)
        grads = torch.zeros(F0 * F1 + F1 * NCLS, device=dev)                      # dense W gradients, one bucket
        seeds_dev = torch.from_numpy(my_seeds.view(np.int32)).to(dev)
        y0_host = torch.empty((BATCH, F1)).pin_memory()
        sizes_pin = torch.zeros((n_steps, 8), dtype=torch.int32).pin_memory()     # LayerMeta of the bottom layer per step
        sizes_top = torch.zeros((n_steps, 8), dtype=torch.int32).pin_memory()
    torch.cuda.synchronize()

    # fixed arena pointers and device-side size addresses of every pipeline slot
    slots = []
    with torch.cuda.stream(st_sample):
        for k in range(P):
            views = (nts._capi.LayerView * 2)()
            check(lib.nb_sampler_sample(sampler._samplers[k], ptr(my_seeds[:BATCH]), BATCH, 0, SEED_SAMPLER + rank, 0,
                                        nts.WeightType.Sum, None, 0xFFFFFFFF, views, 1))
            nd, ns, caps = [C.c_void_p(), C.c_void_p()], [C.c_void_p(), C.c_void_p()], [[C.c_uint32() for _ in range(3)] for _ in range(2)]
            for l in range(2):
                check(lib.nb_sampler_sizes_dev(sampler._samplers[k], l, C.byref(nd[l]), None, C.byref(ns[l]), C.byref(caps[l][0]),
                                               C.byref(caps[l][1]), C.byref(caps[l][2])))
            assert caps[1][2].value <= cap_s1 and caps[0][2].value <= cap_s0
            slots.append(dict(top=views[0], bot=views[1], nd=nd, ns=ns, caps=caps, sampled=torch.cuda.Event(), consumed=torch.cuda.Event()))
    torch.cuda.synchronize()

    ev = lambda: torch.cuda.Event(enable_timing=True)
    kern_ev = {"gather": [], "agg_fwd_602": []}
    st_comm = torch.cuda.Stream(dev)
    comm_box = [None]
    cs_comm = nts.Cuda_Stream(local, st_comm)
    peer_ar = None
    if world > 1 and not args.nccl_allreduce:
        from sample_based_gnn_b200 import dist as nbdist
        peer_ar = nbdist.PeerAllReduce(cs_comm, grads.numel())   # the dense-gradient exchange as one kernel over peer memory

    pending_box = [None]

    def issue_allreduce():
        if pending_box[0] is None:
            return
        st_comm.wait_event(pending_box[0])          # the backward that produced the gradients
        if not args.comm_early:
            after_gather = torch.cuda.Event()
            after_gather.record(st_train)
            st_comm.wait_event(after_gather)
        if peer_ar is not None:
            peer_ar.all_reduce(grads)               # one kernel over NVLink peer memory, enqueued on st_comm
        else:
            with torch.cuda.stream(st_comm):
                dist.all_reduce(grads)
        comm_box[0] = torch.cuda.Event()
        comm_box[0].record(st_comm)
        pending_box[0] = None

    def step_async(i, timed, fused=False):
        """value: inputs resident in HBM, no host synchronisation anywhere (sizes stay on the device). Batch i is sampled on
        the sampling stream into arena i % P while the training stream gathers / aggregates batch i-1."""
        sl = slots[i % P]
        top, bot, nd, ns, caps = sl["top"], sl["bot"], sl["nd"], sl["ns"], sl["caps"]
        st_sample.wait_event(sl["consumed"])
        check(lib.nb_sampler_sample(sampler._samplers[i % P], ptr(seeds_dev[i * BATCH:(i + 1) * BATCH]), BATCH, 1,
                                    SEED_SAMPLER + rank, i, nts.WeightType.Sum, None, 0xFFFFFFFF, None, 0))
        sl["sampled"].record(st_sample)
        st_train.wait_event(sl["sampled"])
        if timed and not fused:
            a, b, c = ev(), ev(), ev()
            a.record(st_train)
        if not fused:
            check(lib.nb_gather_rows_dyn(cs_train._h, ptr(x0), ptr(table), bot.source, ns[1], caps[1][2], F0, PITCH, PITCH))
            if timed:
                b.record(st_train)
            issue_allreduce()   # the previous step's gradient exchange is enqueued here, on its own stream, behind that step's backward:
                                # it overlaps this gather and the aggregation below and has both (~0.23 ms) to absorb rank skew.
                                # --comm-late holds it back until the gather is done: the gather keeps its full bandwidth
                                # (0.82 vs 0.78 of peak at N=2) but at N=8 the exchange then has only the aggregation (~0.1 ms)
                                # to hide in and the step stalls on the slowest rank (0.325 vs 0.295 ms measured).
            check(lib.nb_aggregate_csc_fwd_dyn(cs_train._h, ptr(x0), ptr(y1), bot.edge_weight_forward, bot.row_indices,
                                               bot.column_offset, nd[1], caps[1][0], F0, PITCH, PITCH))
            if timed:
                c.record(st_train)
                kern_ev["gather"].append((a, b))
                kern_ev["agg_fwd_602"].append((b, c))
        else:  # bottom hop aggregated straight from the feature table through the global ids: X0 is never materialised
            issue_allreduce()
            check(lib.nb_aggregate_csc_fwd_dyn(cs_train._h, ptr(table), ptr(y1), bot.edge_weight_forward, bot.sample_ans,
                                               bot.column_offset, nd[1], caps[1][0], F0, PITCH, PITCH))
        if comm_box[0] is not None:
            st_train.wait_event(comm_box[0])
        check(lib.nb_aggregate_csc_fwd_dyn(cs_train._h, ptr(h1), ptr(y0), top.edge_weight_forward, top.row_indices,
                                           top.column_offset, nd[0], caps[0][0], F1, F1, F1))
        check(lib.nb_aggregate_csr_bwd_dyn(cs_train._h, ptr(dy0), ptr(dh1), top.edge_weight_backward, top.row_offset,
                                           top.column_indices, ns[0], caps[0][2], F1, F1, F1))
        check(lib.nb_memcpy_d2h(cs_train._h, ptr(sizes_pin[i]), nd[1].value, 32, 0))
        check(lib.nb_memcpy_d2h(cs_train._h, ptr(sizes_top[i]), nd[0].value, 32, 0))
        sl["consumed"].record(st_train)
        if world > 1:
            # the dense-gradient exchange of this step runs on its own stream behind this step's backward; it is issued by the
            # NEXT step right after its gather launch (issue_allreduce) and the next step's top hop -- the first consumer of the
            # updated weights -- waits for it
            pending_box[0] = torch.cuda.Event()
            pending_box[0].record(st_train)

    api_state = {"issued": -1, "checksum": 0.0}
    y0_ring = [torch.empty((BATCH, F1)).pin_memory() for _ in range(2)]
    y0_done = [torch.cuda.Event(), torch.cuda.Event()]

    def api_issue(i):
        """sample batch i asynchronously on the sampling stream into slot i % 2 (FastSampler pipeline slot, as PIPELINE_NUM=2)"""
        k = i % 2
        st_sample_api.wait_event(api_ev[k]["consumed"])
        fast.work_offset = i * BATCH
        with torch.cuda.stream(st_sample_api):
            fast.sample_gpu_fast(BATCH, ssg_id=k, sync=False)          # stages + uploads the seeds from host memory
        api_ev[k]["sampled"].record(st_sample_api)
        api_state["issued"] = i

    def step_api(i, timed):
        """e2e: the reference-shaped API a user calls -- host seeds in, sizes and the batch's top-layer output back on the
        host every step. Batch i+1 is sampled (pipeline slot (i+1) % 2) while batch i is gathered / aggregated."""
        if api_state["issued"] < i:
            api_issue(i)
        k = i % 2
        sg = fast.wait(k)                                           # host waits for the sizes of batch i only
        st_train.wait_event(api_ev[k]["sampled"])
        t, bt = sg.sampled_sgs
        xx = x0[:bt.src_size, :F0]
        fast.load_feature_gpu(cs_train, sg, xx, table[:, :F0])
        issue_allreduce()                                           # the previous step's gradient exchange, behind this gather
        yy1 = nts.SingleGPUAllSampleGraphOp(sg, 1, cs_train).forward(xx)
        op_top = nts.SingleGPUAllSampleGraphOp(sg, 0, cs_train)
        if comm_box[0] is not None:
            st_train.wait_event(comm_box[0])                        # updated weights before the top hop
        yy0 = op_top.forward(h1[:t.src_size])
        op_top.backward(dy0)
        api_ev[k]["consumed"].record(st_train)
        if world > 1:
            pending_box[0] = torch.cuda.Event()
            pending_box[0].record(st_train)
        if i + 1 < n_steps:
            api_issue(i + 1)                                        # next batch samples while this one gathers / aggregates
        # the step's result goes to one of two pinned host buffers; the host consumes step i-1's while step i runs
        y0_ring[i % 2].copy_(yy0, non_blocking=True)
        y0_done[i % 2].record(st_train)
        if i > 0:
            y0_done[(i - 1) % 2].synchronize()
            api_state["checksum"] += float(y0_ring[(i - 1) % 2][0, 0])
        if i + 1 == n_steps:
            y0_done[i % 2].synchronize()
            api_state["checksum"] += float(y0_ring[i % 2][0, 0])
        sizes_pin[i, 0], sizes_pin[i, 1], sizes_pin[i, 2] = bt.v_size, bt.e_size, bt.src_size
        sizes_top[i, 1] = t.e_size
        del yy1

    clock_box = [None]

    def run(mode, sample_clocks=False):
        step = {"async": step_async, "fused": lambda i, t: step_async(i, t, True), "api": step_api}[mode]
        api_state["issued"] = -1
        comm_box[0] = None
        pending_box[0] = None
        for k in kern_ev.values():
            k.clear()
        with torch.cuda.stream(st_train):
            for i in range(args.warmup):
                step(i, False)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            launches0 = cs_sample.launch_count() + cs_train.launch_count() + cs_sample_api.launch_count() + cs_comm.launch_count()
            clocks = ClockSampler(local) if sample_clocks else None
            if clocks:
                clocks.start()
            t0, t1 = ev(), ev()
            t0.record(st_train)
            for i in range(args.warmup, n_steps):
                step(i, True)
            issue_allreduce()                      # the last step's exchange
            if comm_box[0] is not None:
                st_train.wait_event(comm_box[0])
            t1.record(st_train)   # every batch's sampling is consumed on the training stream, so this closes all streams
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
        launches = cs_sample.launch_count() + cs_train.launch_count() + cs_sample_api.launch_count() + cs_comm.launch_count() - launches0
        return ms, launches, work, {k: sum(a.elapsed_time(b) for a, b in v) / max(len(v), 1) for k, v in kern_ev.items()}

    ms, launches, work, kms = run("async", sample_clocks=True)
    clk = clock_box[0]
    ms_e2e, _, work_e2e, _ = run("api")
    ms_fused, _, work_fused, _ = run("fused")

    def reduce(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    MAX, SUM = (dist.ReduceOp.MAX, dist.ReduceOp.SUM) if world > 1 else (None, None)
    ms_max, ms_e2e_max, ms_fused_max = reduce(ms, MAX), reduce(ms_e2e, MAX), reduce(ms_fused, MAX)
    edges_all, edges_e2e_all, edges_fused_all = reduce(work["edges"], SUM), reduce(work_e2e["edges"], SUM), reduce(work_fused["edges"], SUM)
    launches_all = reduce(launches, SUM)
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
                "kernels": {k: {"gbs": round(x["gbs"], 1), "frac": round(x["gbs"] / peak, 3), "ms": round(x["ms"], 4),
                                "algorithmic_bytes": int(x["algorithmic_bytes"])} for k, x in kernels.items()}}
        traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(traffic_file):
            try:
                tr = json.load(open(traffic_file))
                roof["traffic"] = tr.get(dom)
                roof["traffic_note"] = tr.get("note")
            except Exception:
                pass
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                r, kind, threads = cpu_baseline_run(v, col_off, src, all_seeds, args.cpu_batches, 2)
                val, _ = cpu_metric(r)
                cpu = {"value": val, "unit": "edges/s", "cores": threads, "kind": kind,
                       "sample": f"{r['batches']} mini-batches of {BATCH} seeds of the same workload through "
                                 + ("oracle/_ref/ref_driver (the reference's own OpenMP sample_fast/get_feature/MiniBatchFuseOp), all host threads; "
                                    if kind == "reference" else "oracle/oracle.c (scalar port, reference driver not built); ")
                                 + f"sample/gather/fwd/bwd s = {r['sample_s']:.3f}/{r['gather_s']:.3f}/{r['fwd_s']:.3f}/{r['bwd_s']:.3f}"}
            except Exception as ex:  # the checker must never take the bench down
                cpu = {"value": None, "unit": "edges/s", "cores": 0, "kind": "port", "sample": f"failed: {ex}"}
        line = {"metric": "sampled_edges_per_s", "value": value, "unit": "edges/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_dict(v, e_total, {
                    "per_gpu_batch": BATCH, "pipeline_num": P, "row_pitch_floats": PITCH,
                    "parallelism": (f"dp{world}: seeds sharded contiguously, one bucketed sum all-reduce of the dense grads per step ("
                                    + ("NCCL" if args.nccl_allreduce else "one kernel over NVLink peer memory, nb_peer_allreduce_sum") + ")"
                                    if world > 1 else "single GPU"),
                    "avg_E_per_step": work["edges"] / n, "avg_S1": S1, "avg_E1": E1, "avg_V1": V1,
                    "epoch_ms_est": (ms_max / args.steps) * (all_seeds.size / BATCH / world)}),
                "e2e": {"value": e2e_value, "unit": "edges/s", "h2d_bytes_per_step": BATCH * 4 + 64,
                        "d2h_bytes_per_step": BATCH * F1 * 4 + 3 * 32, "ms_per_step": ms_e2e_max / args.steps,
                        "path": "FastSampler.sample_gpu_fast(slot i+1, async, high-priority stream) || wait(slot i) -> load_feature_gpu -> SingleGPUAllSampleGraphOp fwd/fwd/bwd -> D2H of the output into a 2-deep pinned ring; the host reads step i-1's output while step i runs"},
                "gpu_launches": int(launches_all), "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
                "fused_gather_aggregate": {"value": edges_fused_all / (ms_fused_max * 1e-3), "unit": "edges/s",
                                           "ms_per_step": ms_fused_max / args.steps,
                                           "note": "same results; the bottom hop aggregates straight from the feature table, X0 is never written"}}
        print(json.dumps(line))
    if peer_ar is not None:
        assert not peer_ar.timed_out(), "peer all-reduce: a rank never arrived"
        peer_ar.close()
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
    ap.add_argument("--pipeline", type=int, default=2, help="PIPELINE_NUM: sampler arenas in flight (sampling overlaps training)")
    ap.add_argument("--pitch", type=int, default=608, help="row pitch in floats of the 602-wide tensors (0 = dense 602)")
    ap.add_argument("--cpu-batches", type=int, default=20)
    ap.add_argument("--sample-priority", type=int, default=0, help="CUDA stream priority of the sampling stream (-1 = high)")
    ap.add_argument("--nccl-allreduce", action="store_true", help="exchange the dense gradients with NCCL instead of the peer-memory kernel")
    ap.add_argument("--comm-late", dest="comm_early", action="store_false",
                    help="hold the gradient all-reduce back until the next step's gather has finished (it then overlaps only the aggregation)")
    ap.set_defaults(comm_early=True)
    ap.add_argument("--opt", action="append", default=[], help="name=value passed to nb_set_option (tuning experiments)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        main_reference(args)
    else:
        main_b200(args)
