"""Regenerates tests/golden/*.npz from the reference itself (TEST INFRASTRUCTURE ONLY).

Runs oracle/_ref/ref_driver (the UNMODIFIED reference CPU path, see oracle/ref_driver.cpp and
oracle/Makefile) in `record` mode on cora and on small seeded synthetic graphs and packs what
it wrote into compressed npz fixtures. Needs /root/reference, so it only runs in the build
container; the fixtures it writes are committed and travel to the GPU box.

    python oracle/make_golden.py            # rebuilds every fixture
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import refio  # noqa: E402

REF = os.environ.get("NTS_REFERENCE", "/root/reference")
DRIVER = os.path.join(HERE, "_ref", "ref_driver")
GOLD = os.path.join(ROOT, "tests", "golden")


def synth_graph(V, avg_deg, seed, zero_in_frac=0.1, hub=True):
    """Edge list with unique (src,dst) pairs, some zero-in-degree vertices, power-law-ish in-degree."""
    rng = np.random.default_rng(seed)
    deg = np.minimum((rng.pareto(1.3, V) * avg_deg * 0.4 + 1).astype(np.int64), V - 1)
    deg[rng.random(V) < zero_in_frac] = 0
    pairs = []
    for d in range(V):
        if deg[d] == 0:
            continue
        if hub:  # skewed sources -> heavy dedup
            p = 1.0 / (1.0 + np.arange(V)) ** 0.7
            p /= p.sum()
            src = rng.choice(V, size=deg[d], replace=False, p=p)
        else:
            src = rng.choice(V, size=deg[d], replace=False)
        pairs.append(np.stack([src, np.full(deg[d], d)], 1))
    pairs = np.concatenate(pairs).astype(np.uint32)
    pairs = pairs[rng.permutation(pairs.shape[0])]  # file order != sorted order
    return pairs


def run_record(name, pairs, V, seeds, batch, fanout, F, up_degree=0, weight="sum"):
    with tempfile.TemporaryDirectory() as td:
        ef = os.path.join(td, "g.edge")
        sf = os.path.join(td, "seeds.u32")
        of = os.path.join(td, "rec.bin")
        np.ascontiguousarray(pairs, dtype=np.uint32).tofile(ef)
        np.ascontiguousarray(seeds, dtype=np.uint32).tofile(sf)
        env = dict(os.environ, OMP_NUM_THREADS="1", NTS_ORACLE_CPUS="2")  # 2-1 = 1 worker thread: bit-stable
        subprocess.check_call([DRIVER, "record", ef, str(V), sf, str(batch), ",".join(map(str, fanout)), str(F), of,
                               str(up_degree), weight], env=env, stdout=subprocess.DEVNULL)
        rec = refio.read_record(of)
    flat = dict(meta=np.array([V, batch, F, up_degree, len(fanout), len(rec["batches"]),
                               {"sum": 0, "mean": 1, "none": 2}[weight]], np.int64),
                fanout=np.array(fanout, np.int64), pairs=np.ascontiguousarray(pairs, np.uint32))
    for k, v in rec["graph"].items():
        flat[k] = v
    for bi, b in enumerate(rec["batches"]):
        for k, v in b.items():
            if k != "layers":
                flat[f"b{bi}_{k}"] = v
        for li, l in enumerate(b["layers"]):
            for k, v in l.items():
                flat[f"b{bi}_l{li}_{k}"] = v
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **flat)
    print(name, os.path.getsize(path) // 1024, "KiB", len(rec["batches"]), "batches")


def run_hotness(name, pairs, V, seeds, batch, pipeline, layers):
    """nts::op::preSample through the reference driver: per-super-batch hot-vertex lists + the .bin file it writes."""
    with tempfile.TemporaryDirectory() as td:
        ef, sf, of = os.path.join(td, "g.edge"), os.path.join(td, "seeds.u32"), os.path.join(td, "rec.bin")
        np.ascontiguousarray(pairs, dtype=np.uint32).tofile(ef)
        np.ascontiguousarray(seeds, dtype=np.uint32).tofile(sf)
        # one OpenMP thread: the reference collects the hot ids from a parallel loop in arrival order (core/ntsBaseOp.hpp:384-395)
        env = dict(os.environ, OMP_NUM_THREADS="1", NTS_ORACLE_CPUS="1")
        subprocess.check_call([DRIVER, "hotness", ef, str(V), sf, str(batch), str(pipeline), str(layers), of], env=env,
                              stdout=subprocess.DEVNULL)
        rec = refio.read_record(of)["graph"]
        raw = np.fromfile(os.path.join(td, f"g.pre_sample_b{batch}_fx_p{pipeline}.bin"), dtype=np.uint32)
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, meta=np.array([V, batch, pipeline, layers], np.int64), pairs=np.ascontiguousarray(pairs, np.uint32),
                        seeds=np.ascontiguousarray(seeds, np.uint32), counts=rec["counts"], ids=rec["ids"], top=rec["top"], bin_file=raw)
    print(name, os.path.getsize(path) // 1024, "KiB", rec["counts"].tolist())


def main():
    if not os.path.exists(DRIVER):
        subprocess.check_call(["make", "-C", HERE, "ref"])
    os.makedirs(GOLD, exist_ok=True)
    # 1. cora, the reference's own fixture (BASELINE.json configs[0] shape: batch 1024, fanout 25-10)
    pairs = np.fromfile(os.path.join(REF, "data", "cora.2708.edge.self"), dtype=np.uint32).reshape(-1, 2)
    train = np.array([int(l.split()[0]) for l in open(os.path.join(REF, "data", "cora.mask")) if l.split()[1] == "train"],
                     np.uint32)
    run_record("cora_b1024_f25-10", pairs, 2708, train, 1024, [25, 10], 16)
    # 2. synthetic, hub-heavy, sampling really happens (deg > fanout), ragged last batch, odd F
    g = synth_graph(600, 12, 11)
    rng = np.random.default_rng(5)
    run_record("synth600_f5-3", g, 600, rng.permutation(600)[:250], 100, [5, 3], 7)
    # 3. UP_DEGREE + Mean weights (GraphSAGE-mean path)
    run_record("synth600_updeg_mean", g, 600, rng.permutation(600)[:128], 64, [4, 4], 5, up_degree=1, weight="mean")
    # 4. take-all (fanout -1 on the first layer), three layers
    g2 = synth_graph(300, 6, 23, hub=False)
    run_record("synth300_takeall_3layer", g2, 300, rng.permutation(300)[:90], 45, [-1, 3, 2], 4)
    # 5. Mean weights with global degrees
    run_record("synth300_mean", g2, 300, rng.permutation(300)[:64], 64, [3, 3], 8, weight="mean")
    # 5b. hotness pre-sampling (a11): one hop (2 layers) and two hops (3 layers), ragged last super-batch
    run_hotness("hotness_synth600_l2", g, 600, rng.permutation(600)[:250], 64, 2, 2)
    run_hotness("hotness_synth300_l3", g2, 300, rng.permutation(300)[:100], 16, 3, 3)
    # 6. the shipped hot-vertex list (a11 on-disk layout: u32 counts[] || u32 ids[]), kept as data
    raw = np.fromfile(os.path.join(REF, "data", "cora.2708.edge.pre_sample_b1024_f25-10_p1.bin"), dtype=np.uint32)
    np.savez_compressed(os.path.join(GOLD, "cora_pre_sample_bin.npz"), raw=raw)
    print("pre_sample words", raw.size)


if __name__ == "__main__":
    main()
