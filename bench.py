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
    N > 1: one bucketed sum of the dense weight gradients (602x128 + 128x41 floats) over NVLink peer memory
The dense layers themselves are libtorch and outside the path; the top hop aggregates a resident
synthetic [S_0,128] activation instead. Headline step: the bottom hop reads its input rows straight from
the feature table (lazy load_feature_gpu); "materialized_x0" is the same step with the gather kernel first.
Timing: W warm-up steps, then --windows windows of exactly K steps each (CUDA events on the training
stream, ranks aligned on the device before every window, max over ranks); the median window is reported.

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
REF_GPU_DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_gpu_driver")
CPP_E2E = os.path.join(ROOT, "sample-based-gnn_b200", "lib", "cpp_e2e_bench")
CPP_E2E_ARGS = []          # [pitch, steps, warmup, windows], set by main_b200


def reddit_shaped_graph(scale=1.0):
    """In-edge CSC with a power-law in-degree (mean ~492, min 1) and popularity-skewed sources."""
    cache = os.environ.get("NB_BENCH_GRAPH_CACHE")    # sweeps: the (deterministic) graph is generated once and re-read by later runs
    if cache and scale == 1.0 and os.path.exists(cache + ".npz"):
        z = np.load(cache + ".npz")
        return int(z["v"]), z["col_off"], z["src"]
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
    if cache and scale == 1.0 and int(os.environ.get("RANK", "0")) == 0 and int(os.environ.get("WORLD_SIZE", "1")) == 1:
        np.savez(cache, v=v, col_off=col_off.astype(np.uint32), src=src)
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
    """SM clock and throttle reasons sampled from the first warm-up step of `value` to the end of the last timed window of `e2e`
    (NVML in-process: a background thread at ~0.5 ms period plus one sample by the main thread per timed window, taken while the GPU is
    still executing that window); samples that fall inside a timed window are counted separately."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.mask, self.stop_flag, self.max_sm = index, [], 0, False, None
        self.sm_timed, self.in_window = [], False
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
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                self.sm.append(mhz)
                if self.in_window:
                    self.sm_timed.append(mhz)
                self.mask |= nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                pass
            time.sleep(0.0005)

    def ensure_started(self):
        if not getattr(self, "_started_once", False):
            self._started_once = True
            self.start()

    def mark(self, inside):
        self.in_window = inside

    def sample_now(self):
        """one sample from the calling thread. The timed windows are a few milliseconds long and the host runs ahead of the GPU, so the
        background thread may see none of a window (GIL, NVML latency): the main thread takes one itself right after it has issued a
        window's last step, while the GPU is still executing that window."""
        if self.nv is None:
            return
        try:
            mhz = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
            self.sm.append(mhz)
            self.sm_timed.append(mhz)
            self.mask |= self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            pass

    def summary(self):
        self.stop_flag = True
        reasons = []
        if self.nv is not None:
            nv = self.nv
            for name, bit in (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                              ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap)):
                if self.mask & bit:
                    reasons.append(name)
        if self.is_alive():
            self.join(timeout=1.0)
        use = self.sm_timed or self.sm     # samples taken inside the timed windows; the whole run (warm-up on) if a window was too short
        return {"sm_mhz": int(np.median(use)) if use else None, "sm_max_mhz": self.max_sm, "reasons": reasons,
                "samples": len(self.sm), "samples_in_timed_windows": len(self.sm_timed)}


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


def run_reference_driver(v, col_off, src, seeds, batches, warmup, threads, gpu_box=None):
    """the reference's own CPU path (oracle/_ref/ref_driver); gpu_box (a list) additionally receives the reference's own GPU
    path on the same inputs (oracle/_ref/ref_gpu_driver: its CUDA kernels + cuSPARSE compiled for sm_100) when that binary exists"""
    with tempfile.TemporaryDirectory() as td:
        ef, sf = write_reference_inputs(td, v, col_off, src, seeds)
        env = dict(os.environ, NTS_ORACLE_CPUS=str(threads + 1), OMP_NUM_THREADS=str(threads))
        argv = ["bench", ef, str(v), sf, str(BATCH), ",".join(map(str, FANOUT)), str(F0), str(F1)]
        out = subprocess.run([REF_DRIVER] + argv + [str(batches), str(warmup)], capture_output=True, text=True, env=env, check=True).stdout
        if gpu_box is not None and os.path.exists(REF_GPU_DRIVER):
            try:
                g = subprocess.run([REF_GPU_DRIVER] + argv + ["30", "5"], capture_output=True, text=True, env=env, timeout=600)
                gpu_box.append(json.loads([l for l in g.stdout.splitlines() if l.startswith("{")][-1]))
            except Exception as ex:  # a reported extra, never fatal
                gpu_box.append({"failed": str(ex)[:300]})
        if gpu_box is not None and os.path.exists(CPP_E2E) and CPP_E2E_ARGS:
            # the same e2e loop driven from C++ through the adaptor header (tools/cpp_e2e_bench.cpp), same edge / seed files
            try:
                c = subprocess.run([CPP_E2E, ef, str(v), sf, str(BATCH), ",".join(map(str, FANOUT)), str(F0), str(F1)] + CPP_E2E_ARGS,
                                   capture_output=True, text=True, timeout=600)
                gpu_box.append({"cpp_e2e": json.loads([l for l in c.stdout.splitlines() if l.startswith("{")][-1])})
            except Exception as ex:
                gpu_box.append({"cpp_e2e": {"failed": (str(ex) + " " + (c.stderr[-200:] if "c" in dir() else ""))[:400]}})
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


def cpu_baseline_run(v, col_off, src, seeds, batches, warmup, gpu_box=None):
    """(result dict, kind, cores): the reference's own OpenMP path when its driver is present, else the C port"""
    if os.path.exists(REF_DRIVER):
        threads = os.cpu_count() or 1
        return run_reference_driver(v, col_off, src, seeds, batches, warmup, threads, gpu_box), "reference", threads
    return run_oracle_port(v, col_off, src, seeds, min(batches, 5), min(warmup, 1)), "port", 1


def cpu_metric(r):
    t = r["sample_s"] + r["gather_s"] + r["fwd_s"] + r["bwd_s"]
    return r["edges"] / t, t / max(r["batches"], 1) * 1e3


def config_dict(v, e):
    """identical in both arms (the driver compares the two lines' config); run-specific detail goes under "run" """
    return {"workload": f"Reddit-shaped synthetic graph ({v} vertices, {e} edges, power-law in-degree), GCN_SAMPLE hot path: "
                        f"sample fanout 25-10 + reindex/CSC/CSR/weights + gather F=602 + aggregate fwd 602/128 + bwd 128",
            "batch": BATCH, "fanout": "25-10", "layers": "602-128-41",
            "l2": "inputs larger than L2: each batch reads ~130K distinct random rows (~313 MB) of a 561 MB HBM-resident table"}


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


class Workload:
    """one synthetic graph + feature shape pushed through the hot path (the headline is the Reddit-shaped one)"""

    def __init__(self, name, v, col_off, src, seeds, f0, f1, ncls, fanout, pitch):
        self.name, self.v, self.col_off, self.src, self.seeds = name, v, col_off, src, seeds
        self.F0, self.F1, self.NCLS, self.fanout, self.pitch = f0, f1, ncls, list(fanout), pitch


def run_hot_path(env, args, wl, modes, R, sample_clocks=True):
    """W warm-up steps, then R timed windows of exactly K steps for every arm in `modes` ("fused" always runs). Returns the raw and
    rank-reduced results; everything it allocates is released before it returns."""
    torch, dist, nts, C, lib, check, ptr = env.torch, env.dist, env.nts, env.C, env.lib, env.check, env.ptr
    world, rank, local, dev = env.world, env.rank, env.local, env.dev
    F0, F1, NCLS, FANOUT = wl.F0, wl.F1, wl.NCLS, wl.fanout
    strong = args.scaling == "strong"
    B = BATCH // world if strong else BATCH          # strong: fixed global batch, local batch = BATCH / N (GAT_SAMPLE_ALL_MULTI.hpp:322)
    v, col_off, src, all_seeds = wl.v, wl.col_off, wl.src, wl.seeds
    e_total = int(src.numel() if hasattr(src, "numel") else src.size)
    my_seeds = shard_seeds(all_seeds, rank, world)
    K, W = args.steps, args.warmup
    n_steps = W + R * K                              # per mode: W warm-up steps, then R timed windows of exactly K steps
    reps = -(-n_steps * B // my_seeds.size)
    my_seeds = np.tile(my_seeds, reps)[: n_steps * B]
    P = max(1, args.pipeline)
    PITCH = wl.pitch                                  # row pitch of the F0-wide tensors, in floats

    # --sample-streams NS: pipeline slot k samples on stream k % NS. One sampling stream serialises the batches' sampler graphs;
    # with two, batch i+1's sampling overlaps batch i's (both under the aggregation of earlier batches)
    NS = max(1, min(args.sample_streams, P))
    st_samples = [torch.cuda.Stream(dev, priority=args.sample_priority) for _ in range(NS)]
    cs_samples = [nts.Cuda_Stream(local, s_) for s_ in st_samples]
    st_sample, cs_sample = st_samples[0], cs_samples[0]
    st_train = torch.cuda.Stream(dev, priority=args.train_priority)
    # --agg-stream 1 (default): the bottom hop runs on a stream of its own. Y1 = A X0 involves no weights (GCN / GraphSAGE aggregate first, then
    # apply W), so like sampling and the gather it belongs to the data stage of the pipeline: batch i+1's bottom hop runs beside batch i's
    # weight-dependent chain (top hop forward / backward, and in a trainer the dense layers and the optimizer). Y1 has one buffer per slot.
    st_agg = torch.cuda.Stream(dev) if args.agg_stream else st_train
    cs_train = nts.Cuda_Stream(local, st_train)
    cs_agg = nts.Cuda_Stream(local, st_agg) if args.agg_stream else cs_train
    # e2e path: the host waits for every batch's sampled sizes, so sampling is on its critical path and gets a high-priority
    # stream (its small kernels are scheduled ahead of the resident gather / aggregation blocks of the previous batch). In the
    # device-resident path nothing waits for the sampler, and a normal-priority stream leaves the aggregation undisturbed.
    PA = max(2, args.api_pipeline)                    # e2e: FastSampler pipeline slots (the reference's PIPELINE_NUM); PA - 1 batches are sampled ahead
    NSA = max(1, min(args.sample_streams, PA))
    st_samples_api = [torch.cuda.Stream(dev, priority=-1) for _ in range(NSA)]
    cs_samples_api = [nts.Cuda_Stream(local, s_) for s_ in st_samples_api]
    with torch.cuda.stream(st_train):
        graph = nts.FullyRepGraph(cs_sample, v, column_offset=col_off, row_indices=src)
        # one sampler (arena) per pipeline slot, all on the sampling stream (the reference's PIPELINE_NUM SampledSubgraphs)
        # bottom_csr=False: the bottom hop's backward never runs in the GCN toolkits (core/ntsContext.hpp:443), so its CSR is not built
        sampler = nts.FastSampler(graph, my_seeds, 2, B, FANOUT, pipeline_num=P, cuda_stream=[cs_samples[k_ % NS] for k_ in range(P)], build_csr=True,
                                  bottom_csr=False, rng_seed=SEED_SAMPLER + rank)
        fast = nts.FastSampler(graph, my_seeds, 2, B, FANOUT, pipeline_num=PA, cuda_stream=[cs_samples_api[k_ % NSA] for k_ in range(PA)], build_csr=True,
                               bottom_csr=False, rng_seed=SEED_SAMPLER + rank)
        api_ev = [dict(sampled=torch.cuda.Event(), consumed=torch.cuda.Event(), aggregated=torch.cuda.Event()) for _ in range(PA)]
        gen = torch.Generator(device=dev).manual_seed(0x5EED0002)
        table = torch.empty((v, PITCH), device=dev)                               # HBM-resident feature table
        for a_ in range(0, v, 1 << 22):                                           # filled in chunks: no second table-sized temporary
            b_ = min(v, a_ + (1 << 22))
            if PITCH != F0:
                table[a_:b_, F0:] = 0
            table[a_:b_, :F0] = torch.rand((b_ - a_, F0), generator=gen, device=dev) * 2 - 1
        cap_s1, cap_s0 = min(B * FANOUT[0] * FANOUT[1], v), min(B * FANOUT[0], v)
        x0 = torch.zeros((cap_s1, PITCH), device=dev)
        y1s = [torch.zeros((cap_s0, PITCH), device=dev) for _ in range(P if args.agg_stream else 1)]
        y1 = y1s[0]
        h1 = torch.rand((cap_s0, F1), generator=gen, device=dev)                 # stands in for relu(Y1 W1)
        y0 = torch.empty((B, F1), device=dev)
        dy0 = torch.rand((B, F1), generator=gen, device=dev)
        dh1 = torch.empty((cap_s0, F1), device=dev)
        n_grad = F0 * F1 + F1 * NCLS
        grads_src = torch.rand(n_grad, generator=torch.Generator(device=dev).manual_seed(77 + rank), device=dev) - 0.5
        grads = grads_src.clone()                                                 # dense W gradients, one bucket
        seeds_dev = torch.from_numpy(my_seeds.view(np.int32)).to(dev)
        sizes_both = torch.zeros((n_steps, 16), dtype=torch.int32).pin_memory()   # LayerMeta (n_dst, n_edges, n_src, ...) of both layers per step
        sizes_top, sizes_pin = sizes_both[:, :8], sizes_both[:, 8:]
        sizes_np = sizes_both.numpy()
    torch.cuda.synchronize()

    # fixed arena pointers and device-side size addresses of every pipeline slot
    slots = []
    for k in range(P):
        with torch.cuda.stream(st_samples[k % NS]):
            views = (nts._capi.LayerView * 2)()
            check(lib.nb_sampler_sample(sampler._samplers[k], ptr(my_seeds[:B]), B, 0, SEED_SAMPLER + rank, 0,
                                        nts.WeightType.Sum, None, 0xFFFFFFFF, views, 1))
            nd, ns, caps = [C.c_void_p(), C.c_void_p()], [C.c_void_p(), C.c_void_p()], [[C.c_uint32() for _ in range(3)] for _ in range(2)]
            for l in range(2):
                check(lib.nb_sampler_sizes_dev(sampler._samplers[k], l, C.byref(nd[l]), None, C.byref(ns[l]), C.byref(caps[l][0]),
                                               C.byref(caps[l][1]), C.byref(caps[l][2])))
            assert caps[1][2].value <= cap_s1 and caps[0][2].value <= cap_s0
            slots.append(dict(top=views[0], bot=views[1], nd=nd, ns=ns, caps=caps, sampled=torch.cuda.Event(), consumed=torch.cuda.Event(),
                              aggregated=torch.cuda.Event()))
    torch.cuda.synchronize()

    ev = lambda: torch.cuda.Event(enable_timing=True)
    kern_ev = {"gather": [], "agg_fwd_602": [], "agg_fwd_602_from_table": []}
    # ---- dense-gradient exchange (N > 1) ------------------------------------------------------------------------------------
    # split (default): nb_peer_allreduce_begin right behind the backward IN the training stream (push over NVLink, waits for
    #   nobody), nb_peer_allreduce_end right before the next step's top hop, the first consumer of the updated weights: the next
    #   gather + bottom aggregation (~0.2 ms) absorb rank skew and no second stream competes for SM slots.
    # one / nccl: round 1's placement -- one launch (peer kernel or NCCL) on a communication stream, early or late (--comm-late).
    st_comm = torch.cuda.Stream(dev, priority=-1)
    cs_comm = nts.Cuda_Stream(local, st_comm)
    exchange = args.exchange if world > 1 else "none"
    inline_push = exchange == "split-inline"     # A/B: begin's push kernel in the training stream instead of the library's side stream
    if inline_push:
        exchange = "split"
    peer_ar = None
    if exchange in ("split", "one"):
        from sample_based_gnn_b200 import dist as nbdist
        check(lib.nb_set_option(b"peer_push_side_stream", 0 if inline_push else 1))
        peer_ar = nbdist.PeerAllReduce(cs_train if exchange == "split" else cs_comm, n_grad)
    comm_box, pending_box, open_box = [None], [None], [False]

    def issue_allreduce():      # exchange in ("one", "nccl"): the previous step's exchange, on the communication stream
        if pending_box[0] is None:
            return
        st_comm.wait_event(pending_box[0])          # the backward that produced the gradients
        if not args.comm_early:
            after_gather = torch.cuda.Event()
            after_gather.record(st_train)
            st_comm.wait_event(after_gather)
        if exchange == "one":
            peer_ar.all_reduce(grads)
        else:
            with torch.cuda.stream(st_comm):
                dist.all_reduce(grads)
        comm_box[0] = torch.cuda.Event()
        comm_box[0].record(st_comm)
        pending_box[0] = None

    def exchange_before_consumer():
        if exchange == "split":
            if open_box[0]:
                peer_ar.end(grads)
                open_box[0] = False
        elif comm_box[0] is not None:
            st_train.wait_event(comm_box[0])
            comm_box[0] = None

    def exchange_after_backward():
        if exchange == "split":
            peer_ar.begin(grads)
            open_box[0] = True
        elif exchange in ("one", "nccl"):
            pending_box[0] = torch.cuda.Event()
            pending_box[0].record(st_train)

    def exchange_flush():       # closes the last step's exchange inside the timed region
        if exchange in ("one", "nccl"):
            issue_allreduce()
        exchange_before_consumer()

    tl_box = [None]   # --timeline: per-step events [sample start, sample end, train start, after bottom hop, after top fwd, after top bwd]

    def step_async(i, timed, fused=True):
        """value: inputs resident in HBM, no host synchronisation anywhere (sizes stay on the device). Batch i is sampled on
        the sampling stream into arena i % P while the training stream works on batch i-1. fused: the bottom hop aggregates
        straight from the feature table through the layer's global ids (what load_feature_gpu(lazy=True) + the op do): X0 is
        never written. fused=False materialises X0 first (gather kernel + aggregation over X0), as the reference does."""
        sl = slots[i % P]
        top, bot, nd, ns, caps = sl["top"], sl["bot"], sl["nd"], sl["ns"], sl["caps"]
        st_sample, cs_sample = st_samples[(i % P) % NS], cs_samples[(i % P) % NS]
        st_sample.wait_event(sl["consumed"])
        tl = tl_box[0]
        if tl is not None:
            tl.append([ev() for _ in range(7)])
            tl[-1][0].record(st_sample)
        check(lib.nb_sampler_sample(sampler._samplers[i % P], ptr(seeds_dev[i * B:(i + 1) * B]), B, 1,
                                    SEED_SAMPLER + rank, i, nts.WeightType.Sum, None, 0xFFFFFFFF, None, 0))
        if tl is not None:
            tl[-1][1].record(st_sample)
        # the batch's sizes (LayerMeta of both layers, adjacent: one 64-byte copy) go to the host for the edge count of the metric;
        # on the sampling stream, so the copy engine's latency stays off the training stream
        check(lib.nb_memcpy_d2h(cs_sample._h, ptr(sizes_both[i]), nd[0].value, 64, 0))
        sl["sampled"].record(st_sample)
        st_agg.wait_event(sl["sampled"])        # (implies consumed(i - P): this slot's Y1 buffer is free again)
        y1 = y1s[(i % P) % len(y1s)]
        if tl is not None:
            tl[-1][2].record(st_agg)
        if timed:
            a, b, c = ev(), ev(), ev()
            a.record(st_agg)
        if not fused:
            check(lib.nb_gather_rows_dyn(cs_agg._h, ptr(x0), ptr(table), bot.source, ns[1], caps[1][2], F0, PITCH, PITCH))
            if timed:
                b.record(st_agg)
            if exchange in ("one", "nccl"):
                issue_allreduce()
            check(lib.nb_aggregate_csc_fwd_dyn(cs_agg._h, ptr(x0), ptr(y1), bot.edge_weight_forward, bot.row_indices,
                                               bot.column_offset, nd[1], caps[1][0], F0, PITCH, PITCH))
            if timed:
                c.record(st_agg)
                kern_ev["gather"].append((a, b))
                kern_ev["agg_fwd_602"].append((b, c))
        else:
            if exchange in ("one", "nccl"):
                issue_allreduce()
            # load_feature_gpu + the bottom hop's forward as one kernel: rows come straight from the table through the layer's
            # packed gather index; its hint bit steers L2 (rows the batch reads again are kept, single-use rows are not)
            if args.no_l2_hints:   # A/B: the same rows through the plain global ids, no eviction hints
                check(lib.nb_aggregate_csc_fwd_dyn(cs_agg._h, ptr(table), ptr(y1), bot.edge_weight_forward, bot.sample_ans,
                                                   bot.column_offset, nd[1], caps[1][0], F0, PITCH, PITCH))
            else:
                check(lib.nb_aggregate_gathered_fwd_dyn(cs_agg._h, ptr(table), PITCH, bot.gather_index, ptr(y1), bot.edge_weight_forward,
                                                        bot.column_offset, nd[1], caps[1][0], F0, PITCH))
            if timed:
                b.record(st_agg)
                kern_ev["agg_fwd_602_from_table"].append((a, b))
        if tl is not None:
            tl[-1][3].record(st_agg)
        if st_agg is not st_train:
            sl["aggregated"].record(st_agg)
            st_train.wait_event(sl["aggregated"])   # Y1 of this batch is ready: the weight-dependent chain may start
        if tl is not None:
            tl[-1][6].record(st_train)
        exchange_before_consumer()
        check(lib.nb_aggregate_csc_fwd_dyn(cs_train._h, ptr(h1), ptr(y0), top.edge_weight_forward, top.row_indices,
                                           top.column_offset, nd[0], caps[0][0], F1, F1, F1))
        if tl is not None:
            tl[-1][4].record(st_train)
        check(lib.nb_aggregate_csr_bwd_dyn(cs_train._h, ptr(dy0), ptr(dh1), top.edge_weight_backward, top.row_offset,
                                           top.column_indices, ns[0], caps[0][2], F1, F1, F1))
        if tl is not None:
            tl[-1][5].record(st_train)
        sl["consumed"].record(st_train)
        exchange_after_backward()

    api_state = {"issued": -1, "checksum": 0.0, "wait_s": 0.0}
    y0_ring = [torch.empty((B, F1)).pin_memory() for _ in range(2)]
    y0_np = [t_.numpy() for t_ in y0_ring]             # the host reads the results through numpy views (no tensor indexing per step)
    y0_done = [torch.cuda.Event(), torch.cuda.Event()]

    def api_issue(i):
        """sample batch i asynchronously on the sampling stream into slot i % 2 (FastSampler pipeline slot, as PIPELINE_NUM=2)"""
        k = i % PA
        st_sample_api = st_samples_api[k % NSA]
        st_sample_api.wait_event(api_ev[k]["consumed"])
        fast.work_offset = i * B
        fast.sample_gpu_fast(B, ssg_id=k, sync=False)              # stages + uploads the seeds from host memory; runs on the slot's stream
        api_ev[k]["sampled"].record(st_sample_api)
        api_state["issued"] = i

    def step_api(i, timed):
        """e2e: the reference-shaped API a user calls -- host seeds in, sizes and the batch's top-layer output back on the
        host every step. Batches i+1 .. i+PA-1 are sampled (their own pipeline slots) while batch i is aggregated."""
        while api_state["issued"] < min(i + PA - 2, n_steps - 1):   # (re)fill the pipeline: slots of batches i .. i+PA-2
            api_issue(api_state["issued"] + 1 if api_state["issued"] >= i else i)
        k = i % PA
        tw = time.perf_counter()
        sg = fast.wait(k)                                           # host waits for the sizes of batch i only
        api_state["wait_s"] += time.perf_counter() - tw
        if i + PA - 1 < n_steps:
            api_issue(i + PA - 1)                                   # batch i-1's slot is free again: sample ahead (high-priority streams)
                                                                    # while this batch is aggregated and the host issues its ops
        st_agg.wait_event(api_ev[k]["sampled"])
        t, bt = sg.sampled_sgs
        with torch.cuda.stream(st_agg):     # the weight-free bottom hop: the data stage's stream (Y1 is allocated on it too)
            if args.materialize_x0:
                xx = fast.load_feature_gpu(cs_agg, sg, x0[:bt.src_size, :F0], table[:, :F0])
            else:
                xx = fast.load_feature_gpu(cs_agg, sg, x0[:, :F0], table[:, :F0], lazy=True)   # a promise: nothing is copied
            if exchange in ("one", "nccl"):
                issue_allreduce()
            yy1 = nts.SingleGPUAllSampleGraphOp(sg, 1, cs_agg).forward(xx)   # lazy: aggregates straight from the table
            if st_agg is not st_train:
                api_ev[k]["aggregated"].record(st_agg)
        if st_agg is not st_train:
            st_train.wait_event(api_ev[k]["aggregated"])
        op_top = nts.SingleGPUAllSampleGraphOp(sg, 0, cs_train)
        exchange_before_consumer()                                  # updated weights before the top hop
        yy0 = op_top.forward(h1[:t.src_size])
        op_top.backward(dy0)
        api_ev[k]["consumed"].record(st_train)
        exchange_after_backward()
        # the step's result goes to one of two pinned host buffers; the host consumes step i-1's while step i runs
        y0_ring[i % 2].copy_(yy0, non_blocking=True)
        y0_done[i % 2].record(st_train)
        if i > 0:
            y0_done[(i - 1) % 2].synchronize()
            api_state["checksum"] += float(y0_np[(i - 1) % 2][0, 0])
        sizes_np[i, 8:11] = (bt.v_size, bt.e_size, bt.src_size)    # host bookkeeping of the work done (numpy view of the pinned buffer)
        sizes_np[i, 1] = t.e_size
        del yy1

    barrier_buf = torch.zeros(4, device=dev)

    def align_ranks():
        """host barrier, then every rank meets ON THE DEVICE in the training stream immediately before the window's first event
        (an exchange of four floats through the peer kernels): host skew after the NCCL barrier cannot leak into the window"""
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
            if exchange == "split":
                peer_ar.all_reduce(barrier_buf)

    def run(mode, clocks=None):
        step = {"fused": lambda i, t: step_async(i, t, True), "materialized": lambda i, t: step_async(i, t, False), "api": step_api}[mode]
        api_state["issued"] = -1
        api_state["wait_s"] = 0.0
        comm_box[0], pending_box[0], open_box[0] = None, None, False
        for k in kern_ev.values():
            k.clear()
        wins, issue_ms = [], []
        launches = 0
        with torch.cuda.stream(st_train):
            if clocks:
                clocks.ensure_started()
            for i in range(W):
                step(i, False)
            exchange_flush()
            if peer_ar is not None:
                torch.cuda.synchronize()
                peer_ar.stats(reset=True)
            for w in range(R):
                align_ranks()
                launches0 = sum(c_.launch_count() for c_ in cs_samples + cs_samples_api + ([cs_agg] if cs_agg is not cs_train else [])) + cs_train.launch_count() + cs_comm.launch_count()
                t0, t1 = ev(), ev()
                if clocks:
                    clocks.mark(True)
                h0 = time.perf_counter()
                t0.record(st_train)
                for i in range(W + w * K, W + (w + 1) * K):
                    step(i, True)
                exchange_flush()                   # the last step's exchange
                if mode == "api":                  # the host consumes the last step's output inside the window
                    last = W + (w + 1) * K - 1
                    y0_done[last % 2].synchronize()
                    api_state["checksum"] += float(y0_np[last % 2][0, 0])
                t1.record(st_train)   # every batch's sampling is consumed on the training stream, so this closes all streams
                h1_ = time.perf_counter()
                if clocks:
                    clocks.sample_now()           # the GPU is still inside this window (the host ran ahead)
                torch.cuda.synchronize()
                if clocks:
                    clocks.mark(False)
                launches += sum(c_.launch_count() for c_ in cs_samples + cs_samples_api + ([cs_agg] if cs_agg is not cs_train else [])) + cs_train.launch_count() + cs_comm.launch_count() - launches0
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                wins.append(t0.elapsed_time(t1))
                issue_ms.append((h1_ - h0) * 1e3)
        a, b_ = W, n_steps
        sp, st_ = sizes_pin[a:b_].numpy().astype(np.int64), sizes_top[a:b_].numpy().astype(np.int64)
        per_win = lambda x: [int(x[w * K:(w + 1) * K].sum()) for w in range(R)]
        work = {"edges": per_win(sp[:, 1] + st_[:, 1]), "V1": int(sp[:, 0].sum()) / R, "E1": int(sp[:, 1].sum()) / R, "S1": int(sp[:, 2].sum()) / R}
        kms = {k: sum(x.elapsed_time(y) for x, y in lst) / max(len(lst), 1) for k, lst in kern_ev.items()}
        wait = peer_ar.stats(reset=True) if peer_ar is not None else None
        return dict(ms=wins, issue_ms=issue_ms, launches=launches / R, work=work, kms=kms, wait=wait,
                    host_wait_ms_per_step=api_state["wait_s"] * 1e3 / (W + R * K))

    clocks = ClockSampler(local) if sample_clocks else None
    res = {"fused": run("fused", clocks)}                                           # clocks: sampled through the timed windows of `value` and `e2e`
    res["api"] = run("api", clocks) if "api" in modes else res["fused"]             # --modes: tuning sweeps skip the other arms
    clk = clocks.summary() if clocks else None
    res["materialized"] = run("materialized") if "materialized" in modes else res["fused"]

    # the sampler by itself: one batch after the other on one stream, nothing else on the GPU (its serial latency per batch)
    sampler_alone_us = None
    if sample_clocks:
        torch.cuda.synchronize()
        a_, b_ = ev(), ev()
        n_alone = min(60, n_steps)
        for rep in range(2):                     # first pass warms up
            a_.record(st_samples[0])
            for i in range(n_alone):
                check(lib.nb_sampler_sample(sampler._samplers[0], ptr(seeds_dev[i * B:(i + 1) * B]), B, 1,
                                            SEED_SAMPLER + rank, i, nts.WeightType.Sum, None, 0xFFFFFFFF, None, 0))
            b_.record(st_samples[0])
            torch.cuda.synchronize()
        sampler_alone_us = a_.elapsed_time(b_) / n_alone * 1e3
    timeline = None
    if args.timeline and sample_clocks:     # diagnostic: where a step's time goes on the device (events between the kernels cost ~1 us each: not a bench value)
        tl_box[0] = []
        with torch.cuda.stream(st_train):
            comm_box[0], pending_box[0], open_box[0] = None, None, False
            for i in range(W, W + args.timeline):
                step_async(i, False, True)
            exchange_flush()
        torch.cuda.synchronize()
        tl, tl_box[0] = tl_box[0][8:], None
        ref = tl[0][2]
        T = np.array([[ref.elapsed_time(e) for e in row] for row in tl]) * 1e3      # us; columns: sample start, sample end, bottom start,
        timeline = {"steps": len(tl), "unit": "us, mean over steps",                # bottom end, top fwd end, top bwd end, top chain start
                    "step_period": float(np.diff(T[:, 5]).mean()),
                    "sampler_graph": float((T[:, 1] - T[:, 0]).mean()),
                    "bottom_hop": float((T[:, 3] - T[:, 2]).mean()),
                    "bottom_hop_period": float(np.diff(T[:, 3]).mean()),
                    "bottom_end_to_top_start": float((T[:, 6] - T[:, 3]).mean()),
                    "top_fwd(+exchange end)": float((T[:, 4] - T[:, 6]).mean()),
                    "top_bwd": float((T[:, 5] - T[:, 4]).mean()),
                    "prev_bwd_end_to_top_start": float((T[1:, 6] - T[:-1, 5]).mean()),
                    "sample_end_to_bottom_start": float((T[:, 2] - T[:, 1]).mean())}
        if rank == 0:
            print("timeline " + json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in timeline.items()}), file=sys.stderr, flush=True)

    # ---- exchange correctness, on the exact path the timed region used: bit-identical to the rank-ordered fp32 sum ------------
    exchange_check = None
    if world > 1:
        parts = [torch.rand(n_grad, generator=torch.Generator(device=dev).manual_seed(77 + r), device=dev) - 0.5 for r in range(world)]
        want = parts[0].clone()
        for q in parts[1:]:
            want += q                                   # rank order, fp32: what the peer kernel computes
        with torch.cuda.stream(st_train):
            grads.copy_(grads_src)
            if exchange == "split":
                peer_ar.begin(grads)
                peer_ar.end(grads)
            elif exchange == "one":
                st_comm.wait_stream(st_train)
                peer_ar.all_reduce(grads)
                st_train.wait_stream(st_comm)
            else:
                dist.all_reduce(grads)
        torch.cuda.synchronize()
        same = bool(torch.equal(grads, want))
        close = bool(torch.allclose(grads, want, rtol=1e-5, atol=1e-6))
        flag = torch.tensor([1.0 if same else 0.0, 1.0 if close else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        digest = torch.tensor([float(grads.double().sum().item())], dtype=torch.float64, device=dev)
        dmax, dmin = digest.clone(), digest.clone()
        dist.all_reduce(dmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(dmin, op=dist.ReduceOp.MIN)
        exchange_check = ("bit-identical to the rank-ordered fp32 sum on every rank" if flag[0].item() == 1.0 else
                          "within 1e-5 of the rank-ordered sum (summation order differs)" if flag[1].item() == 1.0 else "MISMATCH")
        if dmax.item() != dmin.item():
            exchange_check += "; RANKS DISAGREE"

    def reduce(x, op):
        if world == 1:
            return list(x)
        t = torch.tensor(x, dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return [float(y) for y in t.tolist()]

    MAX, SUM = (dist.ReduceOp.MAX, dist.ReduceOp.SUM) if world > 1 else (None, None)

    def summarize(r):
        """per window: time = max over ranks, edges = sum over ranks; the reported window is the median one"""
        ms = reduce(r["ms"], MAX)
        edges = reduce(r["work"]["edges"], SUM)
        order = sorted(range(R), key=lambda w: ms[w])
        mid = order[R // 2]
        return {"ms": ms[mid], "edges": edges[mid], "value": edges[mid] / (ms[mid] * 1e-3), "ms_per_step": ms[mid] / K,
                "windows_ms_per_step": [round(x / K, 5) for x in ms],
                "host_issue_ms_per_step": round(float(np.median(reduce(r["issue_ms"], MAX))) / K, 5)}

    sm = {k: summarize(r) for k, r in res.items()}
    launches_all = reduce([res["fused"]["launches"]], SUM)[0]
    wait_all = None
    if peer_ar is not None:
        w_ = res["fused"]["wait"]
        wait_all = {"mean_us_max_over_ranks": round(reduce([w_[1]], MAX)[0], 2), "max_us_over_ranks": round(reduce([w_[2]], MAX)[0], 2),
                    "exchanges_per_rank": w_[0]}

    out = dict(timeline=timeline, sampler_alone_us=sampler_alone_us, res=res, sm=sm, launches_all=launches_all, wait_all=wait_all, exchange_check=exchange_check, clk=clk, exchange=exchange,
               B=B, P=P, PITCH=PITCH, K=K, W=W, R=R, e_total=e_total, n_train=int(all_seeds.size))
    if peer_ar is not None:
        assert not peer_ar.timed_out(), "peer all-reduce: a rank never arrived"
        peer_ar.close()
    assert exchange_check is None or ("MISMATCH" not in exchange_check and "DISAGREE" not in exchange_check), exchange_check
    del sampler, fast, graph, table, x0, y1, y1s, h1, slots
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return out


# ---- the other BASELINE.json configs, measured inside the default run (compact records under "other_configs") ------------------
def power_law_graph_gpu(torch, V, E, seed):
    """in-edge CSC generated on the device (int32 tensors holding u32 values): power-law in-degree, popularity-skewed sources --
    the Reddit-shaped recipe at the products / papers100M sizes, without a host copy"""
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
    co32 = co.to(torch.int32)
    del w, deg, co
    return co32, src


def run_products_config(env, args, peak):
    """configs[2]: GraphSAGE on an ogbn-products-shaped graph (2.45M vertices, ~62M edges, F=100). On a 180 GB part the whole table
    (0.98 GB) is HBM resident, so the "hot cache" is the table itself; the same fused step as the headline, narrow rows."""
    torch = env.torch
    V, E, F = 2449029, 61859140, 100
    co, src = power_law_graph_gpu(torch, V, E, 0xBEEF)
    seeds = np.random.default_rng(3).permutation(V)[:max(8192, int(V * 0.08))].astype(np.uint32)
    wl = Workload("products-shaped", V, co, src, seeds, F, 128, 47, FANOUT, F)
    o = run_hot_path(env, args, wl, ["fused", "materialized"], 3, sample_clocks=False)
    del co, src
    f_, m_ = o["sm"]["fused"], o["sm"]["materialized"]
    wk, K = o["res"]["materialized"]["work"], o["K"]
    S1 = wk["S1"] / K
    g_ms = o["res"]["materialized"]["kms"]["gather"]
    gbs = S1 * (4 + 8 * F) / (g_ms * 1e-3) / 1e9 if g_ms else 0.0
    return {"workload": f"products-shaped synthetic graph ({V} vertices, {o['e_total']} edges), F=100-128-47, fanout 25-10, batch {o['B']} per GPU, feature table HBM resident",
            "value": f_["value"], "unit": "edges/s", "ms_per_step": f_["ms_per_step"], "windows_ms_per_step": f_["windows_ms_per_step"],
            "materialized_x0_ms_per_step": m_["ms_per_step"],
            "gather_rows(F=100)": {"ms": round(g_ms, 4), "gbs": round(gbs, 1), "frac": round(gbs / peak, 3), "rows": int(S1)},
            "exchange_check": o["exchange_check"]}


def run_gat_config(env, args, v, col_off, src, all_seeds, peak):
    """configs[3]: GAT_SAMPLE_ALL_MULTI on the Reddit-shaped graph, data parallel over the run's GPUs (local batch = 1024 / N,
    toolkits/GAT_SAMPLE_ALL_MULTI.hpp:322): merge-src-dst sampling, gather F=602, fused GAT layer (edge softmax inside the
    aggregation) forward + backward on both hops (hidden 128 and 41 classes), dense-gradient exchange. Through the operator API."""
    torch, dist, nts, lib, check = env.torch, env.dist, env.nts, env.lib, env.check
    world, rank, local, dev = env.world, env.rank, env.local, env.dev
    B = max(1, BATCH // world)
    K, W, R = args.steps, max(3, args.warmup), 3
    n_steps = W + R * K
    mine = shard_seeds(all_seeds, rank, world)
    mine = np.tile(mine, -(-n_steps * B // mine.size))[: n_steps * B]
    PO = 3                                            # FastSampler pipeline slots: two batches are sampled ahead, on two streams
    st_ss, st_t = [torch.cuda.Stream(dev, priority=-1) for _ in range(2)], torch.cuda.Stream(dev)
    cs_ss, cs_t = [nts.Cuda_Stream(local, s_) for s_ in st_ss], nts.Cuda_Stream(local, st_t)
    st_s, cs_s = st_ss[0], cs_ss[0]
    st_d = torch.cuda.Stream(dev)                     # data stream: the feature gather of the next batch (no weights involved)
    cs_d = nts.Cuda_Stream(local, st_d)
    H, NC = F1, NCLS
    with torch.cuda.stream(st_t):
        graph = nts.FullyRepGraph(cs_s, v, column_offset=col_off, row_indices=src)
        smp = nts.FastSampler(graph, mine, 2, B, FANOUT, pipeline_num=PO, cuda_stream=[cs_ss[k_ % 2] for k_ in range(PO)], merge_src_dst=True, build_csr=True,
                              rng_seed=SEED_SAMPLER + 100 + rank)
        gen = torch.Generator(device=dev).manual_seed(5)
        table = torch.rand((v, F0), generator=gen, device=dev)
        cap1, cap0 = min(B * 26 * 11, v), min(B * 26, v)
        x0s = [torch.empty((cap1, F0), device=dev) for _ in range(PO)]   # one X0 per slot: the gather is weight-free, it runs on a data stream
        h1, h0 = torch.randn((cap1, H), generator=gen, device=dev), torch.randn((cap0, NC), generator=gen, device=dev)
        att1, att0 = torch.randn(2 * H, generator=gen, device=dev) * 0.3, torch.randn(2 * NC, generator=gen, device=dev) * 0.3
        d1, d0 = torch.randn((cap0, H), generator=gen, device=dev), torch.randn((B, NC), generator=gen, device=dev)
        grads = torch.rand(F0 * H + 2 * H + H * NC + 2 * NC, generator=gen, device=dev)
    torch.cuda.synchronize()
    peer = None
    if world > 1:
        from sample_based_gnn_b200 import dist as nbdist
        peer = nbdist.PeerAllReduce(cs_t, grads.numel())
    ev_s = [dict(sampled=torch.cuda.Event(), consumed=torch.cuda.Event(), gathered=torch.cuda.Event()) for _ in range(PO)]
    state = {"issued": -1, "open": False}
    gat_ev, work = [], []

    def issue(i):
        k = i % PO
        st_s = st_ss[k % 2]
        st_s.wait_event(ev_s[k]["consumed"])
        smp.work_offset = i * B
        with torch.cuda.stream(st_s):
            smp.sample_gpu_fast(B, ssg_id=k, weightType=nts.WeightType.None_, sync=False)
        ev_s[k]["sampled"].record(st_s)
        state["issued"] = i

    def step(i, timed):
        while state["issued"] < min(i + PO - 2, n_steps - 1):
            issue(state["issued"] + 1 if state["issued"] >= i else i)
        k = i % PO
        sg = smp.wait(k)
        st_d.wait_event(ev_s[k]["sampled"])
        top, bot = sg.sampled_sgs
        smp.load_feature_gpu(cs_d, sg, x0s[k][:bot.src_size], table)   # X0 feeds the dense W0, so it is materialised -- on the data stream,
        ev_s[k]["gathered"].record(st_d)                               # beside the previous batch's (weight-dependent, latency-bound) GAT hops
        st_t.wait_event(ev_s[k]["gathered"])
        op1, op0 = nts.GATFusedOp(sg, 1, cs_t), nts.GATFusedOp(sg, 0, cs_t)
        if timed:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st_t)
        o1 = op1.forward(h1[:bot.src_size], att1)                      # hop 1: [S1,128] -> [V1,128]
        if timed:
            b.record(st_t)
            gat_ev.append((a, b, bot.e_size, bot.v_size, bot.src_size))
        if state["open"]:
            peer.end(grads)
            state["open"] = False
        o0 = op0.forward(h0[:top.src_size], att0)                      # hop 0: [S0,41] -> [B,41]
        op0.backward(h0[:top.src_size], att0, d0[:top.v_size])
        op1.backward(h1[:bot.src_size], att1, d1[:bot.v_size])
        ev_s[k]["consumed"].record(st_t)
        if peer is not None:
            peer.begin(grads)
            state["open"] = True
        if i + PO - 1 < n_steps:
            issue(i + PO - 1)
        work.append(top.e_size + bot.e_size)
        del o1, o0

    wins, edges = [], []
    with torch.cuda.stream(st_t):
        for i in range(W):
            step(i, False)
        for w in range(R):
            if state["open"]:
                peer.end(grads); state["open"] = False
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
                peer.all_reduce(torch.zeros(4, device=dev))
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n0 = len(work)
            t0.record(st_t)
            for i in range(W + w * K, W + (w + 1) * K):
                step(i, True)
            if state["open"]:
                peer.end(grads); state["open"] = False
            t1.record(st_t)
            torch.cuda.synchronize()
            wins.append(t0.elapsed_time(t1))
            edges.append(float(sum(work[n0:])))
    t = torch.tensor(wins + edges, dtype=torch.float64, device=dev)
    if world > 1:
        tm, te = t[:R].clone(), t[R:].clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX); dist.all_reduce(te, op=dist.ReduceOp.SUM)
        wins, edges = tm.tolist(), te.tolist()
    mid = sorted(range(R), key=lambda w: wins[w])[R // 2]
    ms = sum(a.elapsed_time(b) for a, b, *_ in gat_ev) / len(gat_ev)
    E1, V1, S1 = (sum(x[k] for x in gat_ev) / len(gat_ev) for k in (2, 3, 4))
    gat_bytes = E1 * 4 + E1 * 4 * H + (V1 + 1) * 4 + 4 * (S1 + V1) + V1 * 4 * H + E1 * 4          # SURVEY section 8(d), fused GAT forward
    gbs = gat_bytes / (ms * 1e-3) / 1e9
    if peer is not None:
        peer.close()
    del smp, graph, table, x0s
    torch.cuda.synchronize(); torch.cuda.empty_cache()
    return {"workload": f"GAT_SAMPLE_ALL_MULTI shape on the Reddit-shaped graph: global batch {B * world} (local {B}), fanout 25-10, merge-src-dst sampling, "
                        f"gather F=602 + fused GAT layer fwd+bwd on both hops (128 / 41) + gradient exchange, operator API with host sizes",
            "value": edges[mid] / (wins[mid] * 1e-3), "unit": "edges/s", "ms_per_step": wins[mid] / K, "scaling": "strong (fixed global batch 1024)",
            "windows_ms_per_step": [round(x / K, 5) for x in wins],
            "roofline": {"bound": "hbm", "kernel": "k_gat_node_scores + k_gat_fwd (hop 1, F'=128)", "ms": round(ms, 4), "achieved": round(gbs, 1),
                         "peak": peak, "unit": "GB/s", "frac": round(gbs / peak, 3), "algorithmic_bytes": int(gat_bytes)}}


def run_papers_config(env, args, peak):
    """configs[4]: GraphSAGE on an ogbn-papers100M-shaped graph (111M vertices, ~1.6B edges, F=128) at FULL size: topology replicated
    (6.9 GB per GPU), the 56.8 GB feature table row-sharded over the run's GPUs in peer-mapped HBM (7.1 GB per GPU at N=8) and read
    over NVLink inside the gather kernel. With 180 GB per GPU every row is HBM resident on some GPU: no host-streamed cold tier is
    needed (that tier -- nb_stage_* -- is exercised by the tests and tools/config_bench.py)."""
    torch, dist, nts = env.torch, env.dist, env.nts
    world, rank, local, dev = env.world, env.rank, env.local, env.dev
    V, E, F, H = 111059956, 1615685872, 128, 128
    free, _ = torch.cuda.mem_get_info()
    need = 6.9e9 * 2.2 + 56.8e9 / world * 2.1 + 4e9
    if free < need:
        return {"skipped": f"needs ~{need / 1e9:.0f} GB of free HBM per GPU at N={world}, {free / 1e9:.0f} GB free"}
    B, K, W, R = BATCH, args.steps, max(3, args.warmup), 3
    PO = 3                                            # FastSampler pipeline slots: two batches are sampled ahead, on two streams
    st_ss, st_t = [torch.cuda.Stream(dev, priority=-1) for _ in range(2)], torch.cuda.Stream(dev)
    cs_ss, cs_t = [nts.Cuda_Stream(local, s_) for s_ in st_ss], nts.Cuda_Stream(local, st_t)
    st_s, cs_s = st_ss[0], cs_ss[0]
    st_d = torch.cuda.Stream(dev)                     # data stream: the feature gather of the next batch (no weights involved)
    cs_d = nts.Cuda_Stream(local, st_d)
    from sample_based_gnn_b200 import dist as nbdist
    with torch.cuda.stream(st_t):
        co, src = power_law_graph_gpu(torch, V, E, 0xFACE)
        e_total = int(src.numel())
        graph = nts.FullyRepGraph(cs_s, V, column_offset=co, row_indices=src)
        del co, src
        torch.cuda.empty_cache()
        n_train = int(V * 0.011)
        train = np.random.default_rng(3).permutation(V)[:n_train].astype(np.uint32)
        mine = shard_seeds(train, rank, world)
        n_steps = W + R * K
        mine = np.tile(mine, -(-n_steps * B // mine.size))[: n_steps * B]
        smp = nts.FastSampler(graph, mine, 2, B, FANOUT, pipeline_num=PO, cuda_stream=[cs_ss[k_ % 2] for k_ in range(PO)], build_csr=True, bottom_csr=False,
                              rng_seed=SEED_SAMPLER + 200 + rank)
        gen = torch.Generator(device=dev).manual_seed(11 + rank)
        n_local = (V - rank + world - 1) // world
        rows = torch.empty((n_local, F), device=dev)
        for a in range(0, n_local, 1 << 22):
            b = min(n_local, a + (1 << 22))
            rows[a:b] = torch.rand((b - a, F), generator=gen, device=dev)
        if world > 1:
            shard = nbdist.ShardedTable(cs_d, rows, V, F)
            del rows
            torch.cuda.empty_cache()
            gather = lambda x, ids, n: shard.gather(x, ids, n)
        else:
            shard = None
            gather = lambda x, ids, n: cs_d.zero_copy_feature_move_gpu(x, rows, ids, F, n, F, F)
        cap1, cap0 = B * 25 * 10, B * 25
        x0s = [torch.empty((cap1, F), device=dev) for _ in range(PO)]    # one X0 per slot: the gather (weight-free) runs on a data stream
        h1, dy0 = torch.rand((cap0, H), device=dev), torch.rand((B, H), device=dev)
        grads = torch.rand(F * H + H * 172, device=dev)
    torch.cuda.synchronize()
    peer = nbdist.PeerAllReduce(cs_t, grads.numel()) if world > 1 else None
    ev_s = [dict(sampled=torch.cuda.Event(), consumed=torch.cuda.Event(), gathered=torch.cuda.Event()) for _ in range(PO)]
    state = {"issued": -1, "open": False}
    g_ev, s_ev, work, rows_n = [], [], [], []

    def issue(i):
        k = i % PO
        st_s = st_ss[k % 2]
        st_s.wait_event(ev_s[k]["consumed"])
        smp.work_offset = i * B
        with torch.cuda.stream(st_s):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st_s)
            smp.sample_gpu_fast(B, ssg_id=k, sync=False)
            b.record(st_s)
            s_ev.append((a, b))
        ev_s[k]["sampled"].record(st_s)
        state["issued"] = i

    def step(i, timed):
        while state["issued"] < min(i + PO - 2, n_steps - 1):
            issue(state["issued"] + 1 if state["issued"] >= i else i)
        k = i % PO
        sg = smp.wait(k)
        st_d.wait_event(ev_s[k]["sampled"])
        top, bot = sg.sampled_sgs
        x0 = x0s[k]
        if timed:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st_d)
        gather(x0[:bot.src_size], bot.dev_source, bot.src_size)          # data stream: beside the previous batch's aggregation
        if timed:
            b.record(st_d)
            g_ev.append((a, b))
            rows_n.append(bot.src_size)
        ev_s[k]["gathered"].record(st_d)
        st_t.wait_event(ev_s[k]["gathered"])
        y1 = nts.SingleGPUAllSampleGraphOp(sg, 1, cs_t).forward(x0[:bot.src_size])
        if state["open"]:
            peer.end(grads); state["open"] = False
        op = nts.SingleGPUAllSampleGraphOp(sg, 0, cs_t)
        op.forward(h1[:top.src_size])
        op.backward(dy0[:top.v_size])
        ev_s[k]["consumed"].record(st_t)
        if peer is not None:
            peer.begin(grads); state["open"] = True
        if i + PO - 1 < n_steps:
            issue(i + PO - 1)
        work.append(top.e_size + bot.e_size)
        del y1

    wins, edges = [], []
    with torch.cuda.stream(st_t):
        for i in range(W):
            step(i, False)
        s_ev.clear()
        for w in range(R):
            if state["open"]:
                peer.end(grads); state["open"] = False
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
                peer.all_reduce(torch.zeros(4, device=dev))
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n0 = len(work)
            t0.record(st_t)
            for i in range(W + w * K, W + (w + 1) * K):
                step(i, True)
            if state["open"]:
                peer.end(grads); state["open"] = False
            t1.record(st_t)
            torch.cuda.synchronize()
            wins.append(t0.elapsed_time(t1))
            edges.append(float(sum(work[n0:])))
    g_ms = sum(a.elapsed_time(b) for a, b in g_ev) / len(g_ev)
    s_ms = sum(a.elapsed_time(b) for a, b in s_ev[:-1]) / max(len(s_ev) - 1, 1)
    rows_avg = sum(rows_n) / len(rows_n)
    stats = torch.tensor(wins + edges + [g_ms, s_ms], dtype=torch.float64, device=dev)
    if world > 1:
        tm, te, tg = stats[:R].clone(), stats[R:2 * R].clone(), stats[2 * R:].clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX); dist.all_reduce(te, op=dist.ReduceOp.SUM); dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        wins, edges, (g_ms, s_ms) = tm.tolist(), te.tolist(), tg.tolist()
    mid = sorted(range(R), key=lambda w: wins[w])[R // 2]
    gbs = rows_avg * (4 + 8 * F) / (g_ms * 1e-3) / 1e9
    remote = (world - 1) / world
    if peer is not None:
        peer.close()
    if shard is not None:
        shard.close()
    del smp, graph, x0s
    torch.cuda.synchronize(); torch.cuda.empty_cache()
    rec = {"workload": f"papers100M-shaped synthetic graph at full size ({V} vertices, {e_total} edges), F=128, fanout 25-10, batch {B} per GPU; "
                       f"topology replicated, 56.8 GB feature table row-sharded over {world} GPU(s) in HBM ({n_local * F * 4 / 1e9:.1f} GB per GPU), "
                       f"peer rows read over NVLink inside the gather kernel; operator API with host sizes",
           "value": edges[mid] / (wins[mid] * 1e-3), "unit": "edges/s", "ms_per_step": wins[mid] / K, "scaling": "weak",
           "windows_ms_per_step": [round(x / K, 5) for x in wins],
           "sampling_ms_per_batch": round(s_ms, 4),
           "gather_rows(F=128)": {"ms": round(g_ms, 4), "rows": int(rows_avg), "gbs_algorithmic_per_gpu": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peak, 3),
                                  "remote_fraction": round(remote, 3)}}
    if world > 1:
        per_peer = rows_avg * remote / (world - 1) * F * 4 / (g_ms * 1e-3) / 1e9
        rec["gather_rows(F=128)"]["nvlink_read_GBps_per_peer"] = round(per_peer, 1)
        rec["gather_rows(F=128)"]["nvlink_read_GBps_total_per_gpu"] = round(per_peer * (world - 1), 1)
        rec["gather_rows(F=128)"]["link_reference"] = "770 GB/s per direction measured between two B200 (profiles/r1_p2p_probe_footprint_sweep.txt)"
    return rec


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
    for kv in args.opt:
        name, value = kv.split("=")
        check(lib.nb_set_option(name.encode(), int(value)))
    from types import SimpleNamespace
    env = SimpleNamespace(torch=torch, dist=dist, nts=nts, C=C, lib=lib, check=check, ptr=ptr, world=world, rank=rank, local=local, dev=dev)
    v, col_off, src = reddit_shaped_graph(args.scale)
    all_seeds = train_seeds(v)
    wl = Workload("reddit-shaped", v, col_off, src, all_seeds, F0, F1, NCLS, FANOUT, args.pitch if args.pitch else F0)
    modes = args.modes.split(",")
    o = run_hot_path(env, args, wl, modes, max(1, args.windows))
    res, sm, K, W, B, P, PITCH, e_total = o["res"], o["sm"], o["K"], o["W"], o["B"], o["P"], o["PITCH"], o["e_total"]
    launches_all, wait_all, exchange_check, clk, exchange = o["launches_all"], o["wait_all"], o["exchange_check"], o["clk"], o["exchange"]
    R = o["R"]
    peak_all = 6650.0
    try:
        peak_all = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0))
    except Exception:
        pass
    other = {}
    if world > 1 and not args.no_other_configs and args.scaling == "weak":
        # strong scaling of the headline step: fixed global batch 1024, local batch 1024 / N (toolkits/GAT_SAMPLE_ALL_MULTI.hpp:322)
        import copy
        a2 = copy.copy(args)
        a2.scaling = "strong"
        try:
            so = run_hot_path(env, a2, wl, ["fused"], 3, sample_clocks=False)
            other["strong scaling of the headline step"] = {
                "global_batch": BATCH, "per_gpu_batch": so["B"], "value": so["sm"]["fused"]["value"], "unit": "edges/s",
                "ms_per_step": so["sm"]["fused"]["ms_per_step"], "windows_ms_per_step": so["sm"]["fused"]["windows_ms_per_step"],
                "steps_per_epoch": int(all_seeds.size // BATCH), "epoch_ms_est": so["sm"]["fused"]["ms_per_step"] * (all_seeds.size // BATCH)}
        except Exception as ex:
            other["strong scaling of the headline step"] = {"failed": f"{type(ex).__name__}: {ex}"[:300]}
    if not args.no_other_configs and args.scale == 1.0:
        for key, fn in (("configs[3] GAT, data parallel", lambda: run_gat_config(env, args, v, col_off, src, all_seeds, peak_all)),
                        ("configs[2] products-shaped", lambda: run_products_config(env, args, peak_all)),
                        ("configs[4] papers100M-shaped", lambda: run_papers_config(env, args, peak_all))):
            try:
                other[key] = fn()
            except Exception as ex:      # an extra record must never take the headline down; all ranks fail or succeed together
                import traceback
                other[key] = {"failed": f"{type(ex).__name__}: {ex}"[:400], "where": traceback.format_exc()[-600:]}
                torch.cuda.synchronize()
                torch.cuda.empty_cache()
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        wk = res["fused"]["work"]
        S1, E1, V1 = wk["S1"] / K, wk["E1"] / K, wk["V1"] / K
        wm = res["materialized"]["work"]
        S1m, E1m, V1m = wm["S1"] / K, wm["E1"] / K, wm["V1"] / K
        bytes_gather = S1m * (4 + 8 * F0)
        bytes_agg = lambda e1, v1: e1 * (8 + 4 * F0) + 4 * (v1 + 1) + 4 * v1 * F0
        kf, km = res["fused"]["kms"], dict(res["materialized"]["kms"])
        if "materialized" not in modes:
            km["gather"] = km["agg_fwd_602"] = 0.0
        kernels = {"segment_reduce_fwd(F=602, rows from the feature table)": {"ms": kf["agg_fwd_602_from_table"], "algorithmic_bytes": bytes_agg(E1, V1)},
                   "gather_rows(F=602)": {"ms": km["gather"], "algorithmic_bytes": bytes_gather},
                   "segment_reduce_fwd(F=602, rows from X0)": {"ms": km["agg_fwd_602"], "algorithmic_bytes": bytes_agg(E1m, V1m)}}
        for x in kernels.values():
            x["gbs"] = x["algorithmic_bytes"] / (x["ms"] * 1e-3) / 1e9 if x["ms"] else 0.0
        dom = "segment_reduce_fwd(F=602, rows from the feature table)"          # the dominant kernel of the headline (fused) step
        roof = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["gbs"] / peak, "traffic": None, "peak_source": peak_src,
                "kernels": {k: {"gbs": round(x["gbs"], 1), "frac": round(x["gbs"] / peak, 3), "ms": round(x["ms"], 4),
                                "algorithmic_bytes": int(x["algorithmic_bytes"])} for k, x in kernels.items()}}
        traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(traffic_file):
            try:
                tr = json.load(open(traffic_file))
                roof["traffic"] = tr.get(dom)
                roof["traffic_note"] = ("NOT measured in this run: " + str(tr.get("note")))
                if roof["traffic"]:
                    roof["dram_side_frac"] = roof["traffic"] / (kernels[dom]["ms"] * 1e-3) / 1e9 / peak
            except Exception:
                pass
        cpu = None
        ref_gpu_box = []
        CPP_E2E_ARGS[:] = [str(PITCH), str(K), str(W), str(R), str(max(2, args.api_pipeline)), str(args.sample_streams), str(args.agg_stream)]
        if world == 1 and not args.no_cpu_baseline:
            try:
                r, kind, threads = cpu_baseline_run(v, col_off, src, all_seeds, args.cpu_batches, 2, ref_gpu_box)
                val, _ = cpu_metric(r)
                cpu = {"value": val, "unit": "edges/s", "cores": threads, "kind": kind,
                       "sample": f"{r['batches']} mini-batches of {BATCH} seeds of the same workload through "
                                 + ("oracle/_ref/ref_driver (the reference's own OpenMP sample_fast/get_feature/MiniBatchFuseOp), all host threads; "
                                    if kind == "reference" else "oracle/oracle.c (scalar port, reference driver not built); ")
                                 + f"sample/gather/fwd/bwd s = {r['sample_s']:.3f}/{r['gather_s']:.3f}/{r['fwd_s']:.3f}/{r['bwd_s']:.3f}"}
            except Exception as ex:  # the checker must never take the bench down
                cpu = {"value": None, "unit": "edges/s", "cores": 0, "kind": "port", "sample": f"failed: {ex}"}
        ref_gpu = None
        cpp_e2e = next((x["cpp_e2e"] for x in ref_gpu_box if "cpp_e2e" in x), None)
        ref_gpu_box = [x for x in ref_gpu_box if "cpp_e2e" not in x]
        if ref_gpu_box:
            g = ref_gpu_box[0]
            ref_gpu = dict(g)
            if "failed" not in g and g.get("batches"):
                nbg = g["batches"]
                ref_gpu["what"] = ("the reference's own GPU path on this box, same graph / seeds / batch (oracle/_ref/ref_gpu_driver = /root/reference/cuda/"
                                   "ntsCUDAGraphOP.cu built for sm_100 + its FastSampler): sample_gpu_fast (host wall clock, it round-trips through the "
                                   "host), zero_copy_feature_move_gpu from its pinned-host table and from an HBM copy, cuSPARSE SpMM fwd F=602 / F=128, "
                                   "bwd F=128; CUDA events, mean ms per batch")
                ref_step = g["sample_ms"] + g["gather_host_table_ms"] + g["spmm_fwd_F0_ms"] + g["spmm_fwd_F1_ms"] + g["spmm_bwd_F1_ms"]
                ref_step_hbm = ref_step - g["gather_host_table_ms"] + g["gather_hbm_table_ms"]
                ref_gpu["serial_step_ms"] = ref_step
                ref_gpu["serial_step_ms_with_hbm_table"] = ref_step_hbm
                ref_gpu["edges_per_s"] = g["edges"] / nbg / (ref_step * 1e-3)
                ref_gpu["edges_per_s_with_hbm_table"] = g["edges"] / nbg / (ref_step_hbm * 1e-3)
                ref_gpu["ours_over_reference_gpu"] = {
                    "step (value / reference with HBM table)": round(sm["fused"]["value"] / ref_gpu["edges_per_s_with_hbm_table"], 2),
                    "gather kernel (HBM table)": round(g["gather_hbm_table_ms"] / km["gather"], 2) if km["gather"] else None,
                    "aggregate fwd F=602 (vs cuSPARSE)": round(g["spmm_fwd_F0_ms"] / km["agg_fwd_602"], 2) if km["agg_fwd_602"] else None}
        f_, a_, m_ = sm["fused"], sm["api"], sm["materialized"]
        ex_name = {"split": ("own kernels over NVLink peer memory: nb_peer_allreduce_begin behind the backward (push, "
                             + ("in the training stream" if args.exchange == "split-inline" else "forked onto the library's side stream")
                             + ") / _end before the next top hop (join + rank-ordered reduce in the training stream)"),
                   "one": "one kernel over NVLink peer memory (nb_peer_allreduce_sum) on a communication stream", "nccl": "NCCL all_reduce on a communication stream",
                   "none": "none (single GPU)"}[exchange]
        line = {"metric": "sampled_edges_per_s", "value": f_["value"], "unit": "edges/s", "n_gpus": world, "steps": K,
                "warmup": W, "ms_per_step": f_["ms_per_step"], "higher_is_better": True, "scaling": args.scaling,
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_dict(v, e_total),
                "run": {"per_gpu_batch": B, "global_batch": B * world, "pipeline_num": P, "sampling_streams": args.sample_streams, "bottom_hop_on_its_own_stream": bool(args.agg_stream), "row_pitch_floats": PITCH,
                        "parallelism": f"dp{world}: seeds sharded contiguously; dense-gradient sum per step = {ex_name}" if world > 1 else "single GPU",
                        "windows": R, "window_rule": "each window times exactly `steps` steps between device-aligned events; the median window is reported",
                        "windows_ms_per_step": f_["windows_ms_per_step"], "host_issue_ms_per_step": f_["host_issue_ms_per_step"],
                        "avg_E_per_step": f_["edges"] / K / world, "avg_S1": S1, "avg_E1": E1, "avg_V1": V1,
                        "epoch_ms_est": f_["ms_per_step"] * (all_seeds.size / (B * world)),
                        "bottom_hop": "aggregated straight from the feature table through the layer's global ids (load_feature_gpu(lazy) + SingleGPUAllSampleGraphOp.forward); bit-identical to gather + aggregate, X0 never written"},
                "collective": ex_name, "exchange_check": exchange_check, "exchange_wait_us": wait_all,
                "e2e": {"value": a_["value"], "unit": "edges/s", "h2d_bytes_per_step": B * 4 + 64,
                        "d2h_bytes_per_step": B * F1 * 4 + 3 * 32, "ms_per_step": a_["ms_per_step"],
                        "windows_ms_per_step": a_["windows_ms_per_step"], "host_issue_ms_per_step": a_["host_issue_ms_per_step"],
                        "host_blocked_in_sampler_wait_ms_per_step": round(res["api"]["host_wait_ms_per_step"], 5),
                        "cpp_host": cpp_e2e,
                        "pipeline_num": max(2, args.api_pipeline),
                        "path": "FastSampler.sample_gpu_fast(slots i+1 .. i+PIPELINE_NUM-1, async, high-priority streams) || wait(slot i) -> [aggregation stream] load_feature_gpu(lazy) -> SingleGPUAllSampleGraphOp fwd/fwd/bwd -> D2H of the output into a 2-deep pinned ring; the host reads step i-1's output while step i runs"},
                "gpu_launches": int(round(launches_all)), "clocks": clk, "roofline": roof, "cpu_baseline": cpu, "reference_gpu": ref_gpu, "other_configs": other,
                "materialized_x0": {"value": m_["value"], "unit": "edges/s", "ms_per_step": m_["ms_per_step"],
                                    "windows_ms_per_step": m_["windows_ms_per_step"],
                                    "note": "same step with X0 materialised first (gather kernel, then aggregation over X0), as the reference's load_feature_gpu does"}}
        if o.get("timeline"):
            line["timeline_diagnostic"] = o["timeline"]
        line["sampler_alone_us_per_batch"] = round(o["sampler_alone_us"], 2) if o.get("sampler_alone_us") else None
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the graph (debug only; the metric is quoted at 1.0)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pipeline", type=int, default=4, help="PIPELINE_NUM: sampler arenas in flight (sampling overlaps training)")
    ap.add_argument("--pitch", type=int, default=608, help="row pitch in floats of the 602-wide tensors (0 = dense 602)")
    ap.add_argument("--cpu-batches", type=int, default=20)
    ap.add_argument("--sample-priority", type=int, default=-1, help="CUDA stream priority of the sampling stream (-1 = high: its small kernels get SM slots ahead of the queued aggregation blocks; 0.184 -> 0.158 ms per step, profiles/r2_sweep_pipeline.txt)")
    ap.add_argument("--exchange", default="one", choices=["split", "split-inline", "one", "nccl"],
                    help="dense-gradient sum at N>1: one (default) = one peer-memory kernel (push + rank-ordered reduce) on a communication stream beside the "
                         "next bottom aggregation; split = push behind the backward + reduce in the training stream before the next top hop "
                         "(the in-line kernel costs ~9 us of the step: 0.156 vs 0.146 ms at N=8, profiles/r2b_scale_matrix.txt); split-inline = "
                         "the same with the push in the training stream too; nccl = NCCL all_reduce")
    ap.add_argument("--modes", default="fused,api,materialized", help="tuning sweeps: run only some arms (a skipped arm repeats the headline's numbers)")
    ap.add_argument("--agg-stream", type=int, default=1, help="1 = the (weight-free) bottom hop runs on its own stream, one Y1 buffer per slot: batch i+1's bottom hop beside batch i's top hop; 0 = everything in the training stream")
    ap.add_argument("--train-priority", type=int, default=-1, help="CUDA stream priority of the training stream (0 = normal, negative = higher). With the bottom hop on its own stream "
                         "the weight-dependent chain (exchange end, top hop forward / backward) is the critical path at N > 1: its few blocks go ahead of the bottom hop's "
                         "(N=2: 0.150 -> 0.133 ms per step; N=1: 0.1265 vs 0.128, profiles/r2c_sweep_n2_train_priority.txt)")
    ap.add_argument("--api-pipeline", type=int, default=4, help="e2e arm: FastSampler pipeline slots (PIPELINE_NUM); PIPELINE_NUM - 1 batches are sampled ahead")
    ap.add_argument("--sample-streams", type=int, default=2, help="sampling streams; pipeline slot k samples on stream k %% NS (one stream serialises the batches' sampler graphs "
                         "and puts them on the step's critical path: 0.168 -> 0.155 ms per step with two, profiles/r2_sweep_sample_streams.txt)")
    ap.add_argument("--timeline", type=int, default=0, help="diagnostic: after the timed arms, N more steps with events between the kernels; prints where a step's time goes")
    ap.add_argument("--windows", type=int, default=5, help="timed windows of exactly --steps steps each; the median window is reported")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="strong: fixed global batch 1024, local batch 1024/N")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the products / GAT / papers100M records (tuning runs)")
    ap.add_argument("--no-l2-hints", action="store_true", help="A/B: gather-fused aggregation without the per-source L2 eviction hints")
    ap.add_argument("--materialize-x0", action="store_true", help="e2e path: gather X0 first instead of the lazy feature handle")
    ap.add_argument("--comm-late", dest="comm_early", action="store_false",
                    help="--exchange one|nccl: hold the all-reduce back until the next step's gather has finished")
    ap.set_defaults(comm_early=True)
    ap.add_argument("--opt", action="append", default=[], help="name=value passed to nb_set_option (tuning experiments)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        main_reference(args)
    else:
        main_b200(args)
