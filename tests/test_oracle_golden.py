"""Pins oracle/oracle.c to the reference: every array the reference's own CPU path produced
(tests/golden, recorded by oracle/make_golden.py through oracle/_ref/ref_driver) must be
reproduced bit for bit by the restatement when it replays the recorded neighbour draws."""
import numpy as np
import pytest

import oracle
from golden_util import LAYER_KEYS, load, names


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("name", names())
def test_graph_build_matches_reference(name):
    g = load(name)
    co, ri = oracle.build_csc(g["pairs"], g["V"])
    assert np.array_equal(co, g["col_off"]) and np.array_equal(ri, g["row_idx"])
    ind, outd = oracle.degrees(g["pairs"], g["V"])
    assert np.array_equal(ind, g["in_deg"]) and np.array_equal(outd, g["out_deg"])


@pytest.mark.parametrize("name", names())
def test_replay_is_bit_exact(name):
    g = load(name)
    wt = {0: 0, 1: 1, 2: None}[g["weight_type"]]
    for b in g["batches"]:
        seeds = b["layers"][0]["destination"]
        lay = oracle.sample_batch(seeds, g["col_off"], g["row_idx"], g["fanout"], g["V"], g["in_deg"], g["out_deg"],
                                  weight_type=wt, up_degree=g["up_degree"],
                                  replay=[l["sample_ans"] for l in b["layers"]])
        for mine, ref in zip(lay, b["layers"]):
            for k in LAYER_KEYS:
                if k in ("e_w_f", "e_w_b") and wt is None:
                    continue
                assert np.array_equal(bits(mine[k]), bits(ref[k])), (name, k)
        if g["up_degree"]:
            assert np.array_equal(lay[-1]["in_deg"], b["deg_in"]) and np.array_equal(lay[-1]["out_deg"], b["deg_out"])
        # gather + aggregate with the degrees the last layer left behind (what the CPU op reads)
        ind = lay[-1]["in_deg"] if g["up_degree"] else g["in_deg"]
        outd = lay[-1]["out_deg"] if g["up_degree"] else g["out_deg"]
        table = oracle.feat(np.arange(g["V"]), np.arange(g["F"]))
        X = oracle.gather_rows(table, lay[-1]["source"])
        assert np.array_equal(bits(X.ravel()), bits(b["X0"]))
        for l in range(g["L"]):
            hop = g["L"] - 1 - l
            L = lay[hop]
            Y = oracle.aggregate_fwd(X, L["column_offset"], L["row_indices"], None, L["destination"], L["source"], ind, outd)
            assert np.array_equal(bits(Y.ravel()), bits(b[f"Y{hop}"])), (name, "Y", hop)
            dY = Y * np.float32(0.5) + np.float32(0.25)
            dX = oracle.aggregate_bwd(dY, L["column_offset"], L["row_indices"], L["source"].size, None,
                                      L["destination"], L["source"], ind, outd)
            assert np.array_equal(bits(dX.ravel()), bits(b[f"dX{hop}"])), (name, "dX", hop)
            X = Y


@pytest.mark.parametrize("name", names())
def test_reference_sampler_obeys_sampling_rule(name):
    """The recorded reference draws themselves: take-all in stored order when deg <= fanout,
    otherwise exactly `fanout` distinct in-neighbours (core/ntsFastSampler.hpp:1028-1048)."""
    g = load(name)
    co, ri = g["col_off"], g["row_idx"]
    for b in g["batches"]:
        for f, l in zip(g["fanout"], b["layers"]):
            for i, d in enumerate(l["destination"]):
                nb = ri[co[d]:co[d + 1]]
                got = l["sample_ans"][l["column_offset"][i]:l["column_offset"][i + 1]]
                if f == -1 or nb.size <= f:
                    assert np.array_equal(got, nb)
                else:
                    # distinct POSITIONS of the column (cora's edge file holds a few duplicate
                    # edges, so an id may repeat up to its multiplicity in the column)
                    assert got.size == f
                    ids, cnt = np.unique(got, return_counts=True)
                    nid, ncnt = np.unique(nb, return_counts=True)
                    assert np.isin(ids, nid).all()
                    assert (cnt <= ncnt[np.searchsorted(nid, ids)]).all()


def test_stored_weights_equal_recomputed_weights():
    g = load("cora_b1024_f25-10")
    b = g["batches"][0]
    l = b["layers"][1]
    X = oracle.feat(np.arange(l["source"].size), np.arange(9))
    y1 = oracle.aggregate_fwd(X, l["column_offset"], l["row_indices"], l["e_w_f"])
    y2 = oracle.aggregate_fwd(X, l["column_offset"], l["row_indices"], None, l["destination"], l["source"], g["in_deg"], g["out_deg"])
    assert np.array_equal(bits(y1), bits(y2))
    dy = y1 + np.float32(1)
    d1 = oracle.aggregate_bwd(dy, l["column_offset"], l["row_indices"], l["source"].size, l["e_w_f"])
    d2 = oracle.aggregate_bwd_csr(dy, l["row_offset"], l["column_indices"], l["e_w_b"])
    assert np.array_equal(bits(d1), bits(d2))


def test_oracle_sampler_distribution_chi_square():
    """The oracle's own sampler is a uniform f-subset sampler: inclusion frequency f/deg."""
    V, deg, f, trials = 64, 40, 7, 4000
    co = np.arange(0, (V + 1) * deg, deg, dtype=np.uint32)
    ri = np.concatenate([np.random.default_rng(v).permutation(V)[:deg] for v in range(V)]).astype(np.uint32)
    dst = np.zeros(trials, np.uint32) + 3
    lco, E = oracle.count_offsets(dst, co, f)
    ans = oracle.sample_layer(dst, lco, co, ri, f, seed=99).reshape(trials, f)
    assert all(np.unique(r).size == f for r in ans)
    nb = ri[co[3]:co[4]]
    cnt = np.array([(ans == v).sum() for v in nb], dtype=np.float64)
    exp = trials * f / deg
    chi2 = ((cnt - exp) ** 2 / exp).sum()
    assert chi2 < 80.0, chi2  # dof 39, p ~ 1e-4


def test_gat_chain_equals_fused_restatement():
    g = load("synth600_f5-3")
    l = g["batches"][0]["layers"][0]
    src, ri, dl = oracle.reindex(l["sample_ans"], l["destination"], g["V"], merge_src_dst=True)
    co = l["column_offset"]
    F = 6
    rng = np.random.default_rng(0)
    H = rng.standard_normal((src.size, F)).astype(np.float32)
    att = rng.standard_normal(2 * F).astype(np.float32)
    out, alpha, pre = oracle.gat_layer_fwd(H, att, co, ri, dl)
    msg = oracle.scatter_src_dst(H, co, ri, dl)
    m = msg @ att
    m = np.where(m > 0, m, np.float32(0.2) * m).astype(np.float32)
    a = oracle.edge_softmax_fwd(m, co)
    out2 = oracle.gather_msg_to_dst(msg[:, :F] * a[:, None], co)
    np.testing.assert_allclose(out, out2, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(alpha, a, rtol=1e-5, atol=1e-7)
    sums = np.add.reduceat(a, co[:-1][np.diff(co) > 0])
    np.testing.assert_allclose(sums, 1.0, rtol=1e-5)
    # backward: chain through the five legacy ops vs the fused restatement
    dout = rng.standard_normal(out.shape).astype(np.float32)
    dmsg_out = oracle.scatter_dst_to_msg(dout, co)                     # d(e_msg_out)
    da = (dmsg_out * msg[:, :F]).sum(1).astype(np.float32)
    dm = oracle.edge_softmax_bwd(da, a, co)
    ds = np.where(pre > 0, dm, np.float32(0.2) * dm).astype(np.float32)
    dmsg = np.zeros_like(msg)
    dmsg[:, :F] = dmsg_out * a[:, None]
    dmsg += ds[:, None] * att[None, :]
    dH = oracle.gather_src_dst(dmsg, co, ri, dl, src.size)
    datt = (msg * ds[:, None]).sum(0)
    dH2, datt2 = oracle.gat_layer_bwd(H, att, dout, pre, alpha, co, ri, dl)
    np.testing.assert_allclose(dH, dH2, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(datt, datt2, rtol=1e-4, atol=1e-4)


def test_pre_sample_bin_layout():
    """a11 on-disk layout: u32 counts[#slots] followed by u32 ids[sum(counts)] (core/ntsBaseOp.hpp:477-541)."""
    import os
    from golden_util import GOLD
    raw = np.load(os.path.join(GOLD, "cora_pre_sample_bin.npz"))["raw"]
    slots = 64
    counts = raw[:slots]
    assert counts.sum() == raw.size - slots
    ids = raw[slots:]
    assert ids.max() < 2708
    off = 0
    for c in counts:
        grp = ids[off:off + c]
        assert np.all(np.diff(grp.astype(np.int64)) > 0)
        off += c


@pytest.mark.parametrize("name", ["hotness_synth600_l2", "hotness_synth300_l3"])
def test_hotness_pre_sampling_matches_reference(name):
    """a11: nts::op::preSample / get_most_neighbor recorded from the reference (1 thread) vs the restatement, and the .bin layout."""
    import os
    from golden_util import GOLD
    z = np.load(os.path.join(GOLD, name + ".npz"))
    V, batch, pipeline, layers = (int(x) for x in z["meta"])
    co, ri = oracle.build_csc(z["pairs"], V)
    counts, ids = oracle.pre_sample(z["seeds"], batch, pipeline, co, ri, V, layers, cache_rate=0.8)  # 0.8: forced by :426-427
    assert np.array_equal(counts, z["counts"]) and np.array_equal(ids, z["ids"])
    assert np.array_equal(oracle.pre_sample_file_pack(counts, ids), z["bin_file"])
    take, sub = oracle.pre_sample_file_unpack(z["bin_file"], counts.size, of_rate=0.5)
    assert np.array_equal(take, (counts * 0.5).astype(np.uint32))
    off = 0
    pos = 0
    for c, t in zip(counts, take):
        assert np.array_equal(sub[pos:pos + t], ids[off:off + t])
        off += int(c)
        pos += int(t)
