"""Loads tests/golden/*.npz (written by oracle/make_golden.py from the reference itself)."""
import glob
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LAYER_KEYS = ["destination", "column_offset", "sample_ans", "source", "row_indices", "row_offset", "column_indices",
              "e_w_f", "e_w_b"]


def names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "*.npz"))
                  if "pre_sample" not in p and "hotness" not in p and not os.path.basename(p).startswith("gat_"))


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    V, batch, F, up_degree, L, nb, wt = (int(x) for x in z["meta"])
    g = dict(V=V, batch=batch, F=F, up_degree=bool(up_degree), L=L, weight_type=wt, fanout=[int(x) for x in z["fanout"]],
             pairs=z["pairs"], col_off=z["g_col_off"], row_idx=z["g_row_idx"], in_deg=z["g_in_deg"],
             out_deg=z["g_out_deg"], batches=[])
    for b in range(nb):
        bd = dict(layers=[])
        for l in range(L):
            ld = {}
            for k in LAYER_KEYS:
                key = f"b{b}_l{l}_{k}"
                if key in z:
                    ld[k] = z[key]
            bd["layers"].append(ld)
        bd["X0"] = z[f"b{b}_X0"]
        if up_degree:
            bd["deg_in"], bd["deg_out"] = z[f"b{b}_deg_in"], z[f"b{b}_deg_out"]
        for hop in range(L):
            bd[f"Y{hop}"] = z[f"b{b}_Y{hop}"]
            bd[f"dX{hop}"] = z[f"b{b}_dX{hop}"]
        g["batches"].append(bd)
    return g
