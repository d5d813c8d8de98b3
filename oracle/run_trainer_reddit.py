"""Runs the reference's own trainer (oracle/_ref/nts_b200 = toolkits/main.cpp linked to libnts_b200.so) on the Reddit-shaped
synthetic graph of bench.py with FEATURE_FILE:random (all-ones features, random labels: core/ntsDataloador.hpp:835-860) and
reports its per-epoch time. TEST/MEASUREMENT INFRASTRUCTURE ONLY.

    python oracle/run_trainer_reddit.py [ALGORITHM] [EPOCHS] [PIPELINE_NUM]
"""
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import bench  # noqa: E402


def main():
    alg = sys.argv[1] if len(sys.argv) > 1 else "GCNSAMPLEALLGPU"
    epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    pipe = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    v, col_off, src = bench.reddit_shaped_graph(1.0)
    with tempfile.TemporaryDirectory() as td:
        ef, _ = bench.write_reference_inputs(td, v, col_off, src, np.zeros(1, np.uint32))
        cfg = os.path.join(td, "reddit.cfg")
        open(cfg, "w").write(f"""ALGORITHM:{alg}
VERTICES:{v}
LAYERS:602-128-41
FANOUT:25-10
BATCH_SIZE:1024
EPOCHS:{epochs}
EDGE_FILE:{ef}
FEATURE_FILE:random
LABEL_FILE:random
MASK_FILE:random
LEARN_RATE:0.001
WEIGHT_DECAY:0.0001
DECAY_RATE:0.97
DECAY_EPOCH:100
DROP_RATE:0.5
PIPELINE_NUM:{pipe}
CACHE_RATE:0.01
FEATURE_CACHE_RATE:0.1
UP_DEGREE:0
PROC_OVERLAP:0
PROC_LOCAL:0
PROC_CUDA:0
PROC_REP:0
LOCK_FREE:1
PUSHDOWN:0
CACHE:0
GPU_NUM:1
""")
        t0 = time.time()
        r = subprocess.run([os.path.join(HERE, "_ref", "nts_b200"), cfg], cwd=os.path.join(HERE, "_ref"), capture_output=True, text=True,
                           env=dict(os.environ, NB_MIRROR_HOST_TABLES=os.environ.get("NB_MIRROR_HOST_TABLES", "1")))
        wall = time.time() - t0
    out = r.stdout + r.stderr
    times = [float(m.group(1)) for m in re.finditer(r"Epoch\[\d+\]:Times\[([0-9.eE+-]+)\(s\)\]", out)]
    print("rc", r.returncode, "wall %.1fs" % wall, "epoch times (s):", times)
    tracing = False
    for l in out.splitlines():
        tracing = tracing or l.startswith("[nts_b200 trace")
        if tracing and (l.startswith("[nts_b200 trace") or l.startswith("nb_") or l.startswith("entry point")):
            print(l)
            continue
        if re.search(r"sample[_ ]time|transfer[_a-z ]*time|train(ing)?[_ ]time|average epoch|run[_ ]time|exec_time|rror", l):
            print(l[-160:])


if __name__ == "__main__":
    main()
