"""one-line digest of a bench.py JSON line read from stdin"""
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = d["roofline"]["kernels"]
print("value %.1fM ms %.4f e2e %.1fM (%.4f ms) fused %.1fM gather %.3f (%.4f ms) agg %.3f (%.4f ms)" % (
    d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, d["e2e"]["ms_per_step"], d["fused_gather_aggregate"]["value"] / 1e6,
    k["gather_rows(F=602)"]["frac"], k["gather_rows(F=602)"]["ms"], k["segment_reduce_fwd(F=602)"]["frac"], k["segment_reduce_fwd(F=602)"]["ms"]))
