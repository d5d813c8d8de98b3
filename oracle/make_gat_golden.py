"""Golden vectors for the GAT edge operators, recorded from the reference's OWN CUDA kernels (TEST INFRASTRUCTURE ONLY).

oracle/_ref/ref_gpu_driver (built by `make -C oracle refgpu` from /root/reference/cuda/ntsCUDAGraphOP.cu + the reference's headers;
oracle/ref_gpu_driver.cpp has no algorithm of its own) runs, on a seeded synthetic sampled layer, the op chain of
toolkits/GAT_SAMPLE_ALL_MULTI.hpp:383-464 with the reference's kernels (cuda/ntsCUDADistKernel.cuh:81-99, 119-133, 174-196, 218-232,
318-388, 440-484) and libtorch for the dense edge NN, forward and backward, and dumps every intermediate.

Needs a GPU (the reference kernels run for real), so it is executed on the B200 box and its outputs are committed:

    gpurun -- python oracle/make_gat_golden.py gpurun_out/gat_golden      # writes gat_<case>.npz there
    cp gpurun_out/gat_golden/gat_*.npz tests/golden/

tests/test_gpu_parity.py::test_gat_against_reference_kernel_records replays them through the C ABI.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from refio import read_record  # noqa: E402

DRIVER = os.path.join(HERE, "_ref", "ref_gpu_driver")
# name: (n_dst, n_src_extra, max_fanout, F, seed, hub)  -- hub: one source that most columns point at (long CSR row)
CASES = {"f8": (200, 300, 12, 8, 1, False), "f41_ragged": (333, 500, 25, 41, 2, False), "f128_hub": (500, 2500, 10, 128, 3, True),
         "f128_top_hop": (1024, 3000, 25, 128, 4, False)}
# per-edge [E,2F] / [E,F] intermediates are kept only for the small cases (they are pure copies / products of kept arrays)
KEEP_EDGE_TENSORS = {"f8"}
EDGE_TENSORS = ["e_msg", "d_e_msg", "e_msg_out", "d_e_msg_out", "cached"]


def layer(n_dst, n_extra, max_f, seed, hub):
    """a merge-src-dst sampled layer: every dst is also a src (dst_local_id), sources are distinct local ids in [0, S)"""
    rng = np.random.default_rng(seed)
    S = n_dst + n_extra
    deg = rng.integers(0, max_f + 1, n_dst)
    deg[rng.random(n_dst) < 0.05] = 0                    # empty columns
    col_off = np.zeros(n_dst + 1, np.uint32)
    np.cumsum(deg, out=col_off[1:])
    row = np.concatenate([rng.choice(S, d, replace=False) for d in deg] + [np.zeros(0, np.int64)]).astype(np.uint32)
    if hub:
        first = col_off[:-1][deg > 0]
        row[first[: len(first) * 3 // 4]] = 7           # local source 7 appears in most columns
    dst_local = rng.permutation(S)[:n_dst].astype(np.uint32)
    return col_off, row, dst_local, S


def write_rec(path, hdr, arrays):
    with open(path, "wb") as f:
        np.asarray(hdr, np.uint32).tofile(f)
        for name, a in arrays.items():
            f.write(name.encode().ljust(16, b"\0"))
            np.asarray([0 if a.dtype == np.uint32 else 1], np.uint32).tofile(f)
            np.asarray([a.size], np.uint64).tofile(f)
            a.tofile(f)
        f.write(b"end".ljust(16, b"\0") + bytes(12))


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    for name, (n_dst, n_extra, max_f, F, seed, hub) in CASES.items():
        co, ri, dl, S = layer(n_dst, n_extra, max_f, seed, hub)
        rng = np.random.default_rng(100 + seed)
        h = rng.standard_normal((S, F)).astype(np.float32)
        att = (rng.standard_normal(2 * F) * 0.3).astype(np.float32)
        dout = rng.standard_normal((n_dst, F)).astype(np.float32)
        with tempfile.TemporaryDirectory() as td:
            fin, fout = os.path.join(td, "in.bin"), os.path.join(td, "out.bin")
            write_rec(fin, [0x4E545352, F, S, 0], {"column_offset": co, "row_indices": ri, "dst_local_id": dl, "h": h.ravel(),
                                                    "att": att, "dout": dout.ravel()})
            subprocess.run([DRIVER, "gat", fin, fout], check=True)
            rec = read_record(fout)["graph"]
        if name not in KEEP_EDGE_TENSORS:
            for k in EDGE_TENSORS:
                rec.pop(k, None)
        np.savez_compressed(os.path.join(out_dir, f"gat_{name}.npz"), column_offset=co, row_indices=ri, dst_local_id=dl, n_src=np.uint32(S),
                            F=np.uint32(F), h=h, att=att, dout=dout, **rec)
        print(name, "E =", ri.size, {k: v.shape for k, v in rec.items()})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(HERE), "gpurun_out", "gat_golden"))
