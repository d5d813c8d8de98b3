"""GPU parity tests: the CUDA path, called through the C ABI (libnts_b200.so), against
 (1) the golden records of the reference's own CPU code (tests/golden, bit-exact on replay),
 (2) the oracle on seeded inputs, (3) size-independent properties at BASELINE.json's sizes.
Tolerance for fp32 aggregation / gradients: 1e-5 relative (north star); integers: bit-exact."""
import numpy as np
import pytest
import torch

import oracle
from golden_util import load, names

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def u32(t):
    return t.detach().cpu().numpy().view(np.uint32)


def f32(t):
    return t.detach().cpu().numpy()


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.fixture(scope="module")
def cs(nts):
    return nts.Cuda_Stream.on_torch_stream(0)  # same stream as torch's ops, like the toolkits


SAMPLER_PATHS = {   # (sampler_fused, sampler_tail bits, sampler_block_threads, sampler_csr_branch)
    "general": (0, 0, 256, 1),        # the default: look-back scans in global memory, any size; CSR kernels on a branch of the graph
    "general-inline": (0, 0, 256, 0),  # the same without the branch
    "general-tails": (0, 3, 256, 1),   # popcount / count scans left to the last block of the sampling / relabel kernels
    "fused": (1, 3, 256, 1),           # small-shape kernels (prefix sums per block in shared memory) + relabel tail
    "fused-512": (1, 2, 512, 0),
    "fused-notail": (1, 0, 512, 1),
}


def set_sampler_path(nts, name):
    lib, check = nts._capi.lib(), nts._capi.check
    fused, tail, block, branch = SAMPLER_PATHS[name]
    check(lib.nb_set_option(b"sampler_fused", fused))
    check(lib.nb_set_option(b"sampler_tail", tail))
    check(lib.nb_set_option(b"sampler_block_threads", block))
    check(lib.nb_set_option(b"sampler_csr_branch", branch))


@pytest.fixture(params=list(SAMPLER_PATHS))
def sampler_path(nts, request):
    """every sampler pipeline: the small-shape kernels in their variants and the general one"""
    set_sampler_path(nts, request.param)
    yield request.param
    set_sampler_path(nts, "general")


def make_graph(nts, cs, V, avg_deg, seed, unique=True, max_deg=None):
    rng = np.random.default_rng(seed)
    deg = np.minimum((rng.pareto(1.5, V) * avg_deg * 0.5).astype(np.int64), max_deg or V - 1)
    deg[rng.random(V) < 0.05] = 0
    cols = []
    for v in range(V):
        if deg[v]:
            src = rng.choice(V, deg[v], replace=False) if unique else rng.integers(0, V, deg[v])
            cols.append(np.stack([src, np.full(deg[v], v)], 1))
    pairs = np.concatenate(cols).astype(np.uint32)
    pairs = pairs[rng.permutation(len(pairs))]
    return pairs, nts.FullyRepGraph(cs, V, edge_pairs=pairs)


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", names())
def test_replay_reference_records_bit_exact(nts, cs, name, sampler_path):
    """north star check 1: replaying the reference's recorded neighbour sets gives bit-exact
    subgraph indices, CSC, CSR and weights; gather / aggregate / gradients follow."""
    g = load(name)
    graph = nts.FullyRepGraph(cs, g["V"], column_offset=g["col_off"], row_indices=g["row_idx"], in_degree=g["in_deg"],
                              out_degree=g["out_deg"])
    wt = {0: nts.WeightType.Sum, 1: nts.WeightType.Mean, 2: nts.WeightType.None_}[g["weight_type"]]
    table_np = oracle.feat(np.arange(g["V"]), np.arange(g["F"]))
    table = torch.from_numpy(table_np).cuda()
    for b in g["batches"]:
        seeds = b["layers"][0]["destination"]
        sampler = nts.FastSampler(graph, seeds, g["L"], max(len(seeds), 1), g["fanout"], cuda_stream=cs,
                                  up_degree=g["up_degree"], build_csr=True)
        sg = sampler.replay(seeds, [l["sample_ans"] for l in b["layers"]], wt)
        for mine, ref in zip(sg.sampled_sgs, b["layers"]):
            assert mine.v_size == ref["destination"].size and mine.e_size == ref["sample_ans"].size
            assert mine.src_size == ref["source"].size
            assert np.array_equal(u32(mine.dev_destination), ref["destination"])
            assert np.array_equal(u32(mine.dev_column_offset), ref["column_offset"])
            assert np.array_equal(u32(mine.dev_source), ref["source"])
            assert np.array_equal(u32(mine.dev_row_indices), ref["row_indices"])
            assert np.array_equal(u32(mine.dev_row_offset), ref["row_offset"])
            assert np.array_equal(u32(mine.dev_column_indices), ref["column_indices"])
            if g["weight_type"] != 2:
                assert np.array_equal(u32(mine.dev_edge_weight_forward), bits(ref["e_w_f"]))
                assert np.array_equal(u32(mine.dev_edge_weight_backward), bits(ref["e_w_b"]))
        # gather: bit-exact
        bottom = sg.sampled_sgs[-1]
        x0 = torch.empty((bottom.src_size, g["F"]), device="cuda")
        sampler.load_feature_gpu(cs, sg, x0, table)
        assert np.array_equal(bits(f32(x0).ravel()), bits(b["X0"]))
        if g["up_degree"] or g["weight_type"] != 0:
            continue  # the CPU op recomputes Sum weights from whatever degrees are current; covered by the Sum fixtures
        X = x0
        for l in range(g["L"]):
            hop = g["L"] - 1 - l
            op = nts.SingleGPUAllSampleGraphOp(sg, hop, cs)
            Y = op.forward(X)
            ref_y = b[f"Y{hop}"].reshape(Y.shape)
            np.testing.assert_allclose(f32(Y), ref_y, rtol=RTOL, atol=1e-7)
            assert np.array_equal(bits(f32(Y)), bits(ref_y)), "forward is expected to be bit-exact (same order, mul+add)"
            dY = torch.from_numpy(ref_y * np.float32(0.5) + np.float32(0.25)).cuda()
            dX = op.backward(dY)
            ref_dx = b[f"dX{hop}"].reshape(dX.shape)
            np.testing.assert_allclose(f32(dX), ref_dx, rtol=RTOL, atol=1e-7)
            # the CSC push (reference GPU backward) agrees too, within tolerance (atomic order)
            dX2 = torch.empty_like(dX)
            lay = sg.sampled_sgs[hop]
            cs.Push_From_Dst_To_Src_Spmm(dY, dX2, lay.dev_e_w(), lay.dev_r_i(), lay.dev_c_o(), lay.src_size, 0, 0, 0, 0,
                                         lay.e_size, lay.v_size, dY.shape[1], True, False)
            np.testing.assert_allclose(f32(dX2), ref_dx, rtol=1e-4, atol=1e-6)
            X = Y


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fanout,build_csr,merge,up", [([25, 10], True, False, False), ([5, 5, 5], True, True, False),
                                                        ([40, 3], True, False, True), ([-1, 4], False, False, False)])
def test_gpu_sampler_pipeline_matches_oracle_on_its_own_draws(nts, cs, fanout, build_csr, merge, up, sampler_path):
    V = 20000
    pairs, graph = make_graph(nts, cs, V, 30, seed=7)
    co, ri = oracle.build_csc(pairs, V)
    ind, outd = oracle.degrees(pairs, V)
    rng = np.random.default_rng(1)
    seeds = rng.permutation(V)[:777].astype(np.uint32)
    sampler = nts.FastSampler(graph, seeds, len(fanout), 1024, fanout, cuda_stream=cs, merge_src_dst=merge, up_degree=up,
                              build_csr=build_csr)
    sg = sampler.sample_gpu_fast(1024)
    ans = [u32(l.dev_sample_ans) for l in sg.sampled_sgs]
    ref = oracle.sample_batch(seeds, co, ri, fanout, V, ind, outd, up_degree=up, merge_src_dst=merge, replay=ans)
    for i, (l, r) in enumerate(zip(sg.sampled_sgs, ref)):
        assert np.array_equal(u32(l.dev_column_offset), r["column_offset"]), i
        assert np.array_equal(u32(l.dev_source), r["source"]), i
        assert np.array_equal(u32(l.dev_row_indices), r["row_indices"]), i
        assert np.array_equal(u32(l.dev_edge_weight_forward), bits(r["e_w_f"])), i
        if build_csr:
            assert np.array_equal(u32(l.dev_row_offset), r["row_offset"]), i
            assert np.array_equal(u32(l.dev_column_indices), r["column_indices"]), i
            assert np.array_equal(u32(l.dev_edge_weight_backward), bits(r["e_w_b"])), i
            c2c = u32(l.dev_csr_to_csc)
            assert np.array_equal(np.sort(c2c), np.arange(l.e_size, dtype=np.uint32))
        use = np.bincount(r["row_indices"], minlength=l.src_size).astype(np.uint32)
        assert np.array_equal(u32(l.dev_source_use_count), use), i            # per-source use counts (the CSR row lengths)
        if i == len(fanout) - 1:                                               # bottom layer: packed gather index = id | hint bit
            gi = u32(l.dev_gather_index)
            assert np.array_equal(gi & 0x7FFFFFFF, ans[i]) and np.array_equal(gi >> 31, (use[r["row_indices"]] >= 3).astype(np.uint32)), i   # gather_keep_min_uses = 3
        else:
            assert l.dev_gather_index is None
        if merge:
            assert np.array_equal(u32(l.dev_dst_local_id), r["dst_local_id"]), i
            s2d = u32(l.dev_src_to_dst)
            dl = r["dst_local_id"]
            assert np.array_equal(s2d[dl], np.arange(l.v_size, dtype=np.uint32))
            assert (s2d != 0xFFFFFFFF).sum() == l.v_size
        # sampling rule (core/ntsFastSampler.hpp:1028-1048)
        lco = r["column_offset"]
        for j, d in enumerate(r["destination"]):
            nb = ri[co[d]:co[d + 1]]
            got = ans[i][lco[j]:lco[j + 1]]
            f = fanout[i]
            if f == -1 or nb.size <= f:
                assert np.array_equal(got, nb)
            else:
                assert got.size == f and np.unique(got).size == f and np.isin(got, nb).all()


@pytest.mark.parametrize("fanout,merge", [([25, 10], False), ([5, 5, 5], True)])
def test_two_level_dedup_bitmap_on_a_large_graph(nts, cs, fanout, merge):
    """The two-level dedup bitmap (O(|V|/1024 + S + E) per layer, chosen by density for very large graphs): every array must still equal the oracle's,
    over several batches on the same sampler (the bitmaps are cleared through level 1, including the odd-layer-count case)."""
    V, E = 1_500_000, 6_000_000
    rng = np.random.default_rng(21)
    pairs = np.stack([rng.integers(0, V, E), rng.integers(0, V, E)], 1).astype(np.uint32)
    graph = nts.FullyRepGraph(cs, V, edge_pairs=pairs)
    co, ri = oracle.build_csc(pairs, V)
    ind, outd = oracle.degrees(pairs, V)
    seeds_all = rng.permutation(V)[:3 * 1024].astype(np.uint32)
    nts._capi.check(nts._capi.lib().nb_set_option(b"sampler_two_level", 1))       # by density this size would still take the flat layout
    nts._capi.check(nts._capi.lib().nb_set_option(b"sampler_fused", 0))           # the small-shape relabel keeps its ranks in shared memory
    try:
        sampler = nts.FastSampler(graph, seeds_all, len(fanout), 1024, fanout, cuda_stream=cs, merge_src_dst=merge, build_csr=True)
    finally:
        nts._capi.check(nts._capi.lib().nb_set_option(b"sampler_two_level", -1))
        nts._capi.check(nts._capi.lib().nb_set_option(b"sampler_fused", 1))
    for b in range(3):
        seeds = seeds_all[b * 1024:(b + 1) * 1024]
        sg = sampler.sample_gpu_fast(1024)
        ans = [u32(l.dev_sample_ans) for l in sg.sampled_sgs]
        ref = oracle.sample_batch(seeds, co, ri, fanout, V, ind, outd, merge_src_dst=merge, replay=ans)
        for i, (l, r) in enumerate(zip(sg.sampled_sgs, ref)):
            assert l.src_size == r["source"].size, (b, i)
            assert np.array_equal(u32(l.dev_source), r["source"]), (b, i)
            assert np.array_equal(u32(l.dev_column_offset), r["column_offset"]), (b, i)
            assert np.array_equal(u32(l.dev_row_indices), r["row_indices"]), (b, i)
            assert np.array_equal(u32(l.dev_row_offset), r["row_offset"]) and np.array_equal(u32(l.dev_column_indices), r["column_indices"]), (b, i)
            assert np.array_equal(u32(l.dev_edge_weight_forward), bits(r["e_w_f"])), (b, i)
            if merge:
                assert np.array_equal(u32(l.dev_dst_local_id), r["dst_local_id"]), (b, i)


def test_sampler_paths_agree_bit_for_bit(nts, cs):
    """the small-shape kernels and the general pipeline: same RNG counters -> same draws -> every array identical; hub rows in the CSR"""
    lib, check = nts._capi.lib(), nts._capi.check
    V = 30000
    pairs, graph = make_graph(nts, cs, V, 35, seed=11)
    hub = np.stack([np.full(6000, 17, np.uint32), np.random.default_rng(5).permutation(V)[:6000].astype(np.uint32)], 1)
    graph = nts.FullyRepGraph(cs, V, edge_pairs=np.concatenate([pairs, hub]))        # vertex 17 is a source of 6000 columns
    seeds = np.random.default_rng(2).permutation(V)[:1024].astype(np.uint32)
    out = {}
    for fused in SAMPLER_PATHS:
        set_sampler_path(nts, fused)
        for fanout, merge, up in (([25, 10], False, False), ([40, 4, 3], True, False), ([6, 6], False, True), ([3, 2, 2], False, False)):
            sm = nts.FastSampler(graph, seeds, len(fanout), 1024, fanout, cuda_stream=cs, merge_src_dst=merge, up_degree=up, build_csr=True)
            for _ in range(3):           # the third replay of the captured graph: counters and bitmaps must have re-armed themselves
                sm.work_offset = 0
                sg = sm.sample_gpu_fast(1024)
            arrs = []
            for l in sg.sampled_sgs:
                arrs += [u32(l.dev_column_offset), u32(l.dev_sample_ans), u32(l.dev_source), u32(l.dev_row_indices), u32(l.dev_row_offset),
                         u32(l.dev_column_indices), u32(l.dev_csr_to_csc), u32(l.dev_edge_weight_forward), u32(l.dev_edge_weight_backward)]
                if merge:
                    arrs += [u32(l.dev_dst_local_id), u32(l.dev_src_to_dst)]
            out[(fused, tuple(fanout))] = arrs
    set_sampler_path(nts, "general")
    for (fused, fan), arrs in out.items():
        if fused != "general":
            other = out[("general", fan)]
            assert len(arrs) == len(other)
            for k, (a, b) in enumerate(zip(arrs, other)):
                assert np.array_equal(a, b), (fused, fan, k)


def test_gpu_sampler_is_reproducible_and_counter_based(nts, cs):
    V = 5000
    pairs, graph = make_graph(nts, cs, V, 40, seed=3)
    seeds = np.arange(512, dtype=np.uint32)
    a = nts.FastSampler(graph, seeds, 2, 512, [10, 5], cuda_stream=cs, rng_seed=123)
    b = nts.FastSampler(graph, seeds, 2, 512, [10, 5], cuda_stream=cs, rng_seed=123)
    sa, sb = a.sample_gpu_fast(512), b.sample_gpu_fast(512)
    assert all(np.array_equal(u32(x.dev_sample_ans), u32(y.dev_sample_ans)) for x, y in zip(sa.sampled_sgs, sb.sampled_sgs))
    a.restart()
    sa2 = a.sample_gpu_fast(512)  # batch counter advanced -> different draws
    assert not np.array_equal(u32(sa2.sampled_sgs[0].dev_sample_ans), u32(sb.sampled_sgs[0].dev_sample_ans))
    c = nts.FastSampler(graph, seeds, 2, 512, [10, 5], cuda_stream=cs, rng_seed=124)
    sc = c.sample_gpu_fast(512)
    assert not np.array_equal(u32(sc.sampled_sgs[0].dev_sample_ans), u32(sb.sampled_sgs[0].dev_sample_ans))


@pytest.mark.parametrize("deg,f", [(40, 7), (26, 25), (100, 40), (700, 10), (33, 32)])
def test_gpu_sampler_chi_square_uniform_inclusion(nts, cs, deg, f):
    """north star check 2: distribution equivalence. Every in-neighbour of a vertex with deg > f is
    included with probability f/deg (uniform f-subsets); chi-square over the inclusion counts."""
    V, trials = 1024, 4096
    rng = np.random.default_rng(deg * 131 + f)
    nbrs = rng.permutation(V)[:deg].astype(np.uint32)
    # vertex 0 has the column under test; every other vertex has one self loop
    pairs = np.concatenate([np.stack([nbrs, np.zeros(deg, np.uint32)], 1),
                            np.stack([np.arange(1, V), np.arange(1, V)], 1).astype(np.uint32)])
    graph = nts.FullyRepGraph(cs, V, edge_pairs=pairs)
    seeds = np.zeros(1, np.uint32)
    sampler = nts.FastSampler(graph, np.zeros(trials, np.uint32), 1, 1, [f], cuda_stream=cs, build_csr=False, rng_seed=99)
    counts = np.zeros(V, np.int64)
    pair_first = np.zeros(V, np.int64)
    for t in range(trials):
        sg = sampler.sample_gpu_fast(1)
        got = u32(sg.sampled_sgs[0].dev_sample_ans)
        assert got.size == f and np.unique(got).size == f
        counts[got] += 1
        pair_first[got[0]] += 1
    assert counts[np.setdiff1d(np.arange(V), nbrs)].sum() == 0
    exp = trials * f / deg
    chi2 = ((counts[nbrs] - exp) ** 2 / exp).sum() / (1 - f / deg)  # hypergeometric variance correction
    dof = deg - 1
    assert chi2 < dof + 5.0 * np.sqrt(2 * dof), (chi2, dof)
    # the slot-0 element is itself uniform over the neighbours
    exp0 = trials / deg
    chi0 = ((pair_first[nbrs] - exp0) ** 2 / exp0).sum()
    assert chi0 < dof + 5.0 * np.sqrt(2 * dof), (chi0, dof)


def test_gpu_sampler_matches_oracle_sampler_distribution(nts, cs):
    """Same statistic from the oracle's sampler and the GPU sampler: two-sample chi-square."""
    V, deg, f, trials = 256, 60, 9, 3000
    rng = np.random.default_rng(5)
    nbrs = rng.permutation(V)[:deg].astype(np.uint32)
    pairs = np.stack([nbrs, np.zeros(deg, np.uint32)], 1)
    graph = nts.FullyRepGraph(cs, V, edge_pairs=pairs)
    co, ri = oracle.build_csc(pairs, V)
    sampler = nts.FastSampler(graph, np.zeros(trials, np.uint32), 1, trials, [f], cuda_stream=cs, build_csr=False)
    sg = sampler.sample_gpu_fast(trials)  # the same dst repeated: independent draws per slot
    got = u32(sg.sampled_sgs[0].dev_sample_ans).reshape(trials, f)
    lco, _ = oracle.count_offsets(np.zeros(trials, np.uint32), co, f)
    ora = oracle.sample_layer(np.zeros(trials, np.uint32), lco, co, ri, f, seed=17).reshape(trials, f)
    a = np.array([(got == v).sum() for v in nbrs], np.float64)
    b = np.array([(ora == v).sum() for v in nbrs], np.float64)
    chi2 = ((a - b) ** 2 / (a + b)).sum()
    assert chi2 < (deg - 1) + 5.0 * np.sqrt(2 * (deg - 1)), chi2


def test_omit_hot_vertices_bottom_layer(nts, cs):
    V = 4000
    pairs, graph = make_graph(nts, cs, V, 20, seed=11)
    co, ri = oracle.build_csc(pairs, V)
    rng = np.random.default_rng(2)
    seeds = rng.permutation(V)[:300].astype(np.uint32)
    flag = np.full(V, 0xFFFFFFFF, np.uint32)
    hot = rng.permutation(V)[:800]
    flag[hot] = 3
    dflag = torch.from_numpy(flag.view(np.int32)).cuda()
    for value in (3, 0xFFFFFFFF):
        sampler = nts.FastSampler(graph, seeds, 2, 300, [6, 4], cuda_stream=cs)
        sg = sampler.sample_gpu_fast_omit(300, dflag, value)
        top, bottom = sg.sampled_sgs
        ref0, _ = oracle.count_offsets(seeds, co, 6)
        assert np.array_equal(u32(top.dev_column_offset), ref0)  # upper layers are not omitted
        ref1, _ = oracle.count_offsets(u32(bottom.dev_destination), co, 4, skip=flag, skip_value=value)
        assert np.array_equal(u32(bottom.dev_column_offset), ref1)
        lens = np.diff(ref1)
        assert (lens[np.isin(u32(bottom.dev_destination), hot)] == 0).all()


def test_capacity_is_checked_not_asserted(nts, cs):
    V = 3000
    pairs, graph = make_graph(nts, cs, V, 20, seed=13)
    sampler = nts.FastSampler(graph, np.arange(100, dtype=np.uint32), 1, 10, [3], cuda_stream=cs)
    with pytest.raises(nts.NtsError):
        sampler.sample_gpu_fast(100)  # more seeds than max_batch
    with pytest.raises(nts.NtsError):
        nts.FastSampler(graph, np.arange(10, dtype=np.uint32), 1, 10, [600], cuda_stream=cs)  # unsupported fanout


def test_empty_and_degenerate_batches(nts, cs):
    V = 64
    pairs = np.array([[1, 0], [2, 0], [3, 1]], np.uint32)  # most vertices have no in-edges
    graph = nts.FullyRepGraph(cs, V, edge_pairs=pairs)
    sampler = nts.FastSampler(graph, np.array([5, 6, 7, 0], np.uint32), 2, 8, [2, 2], cuda_stream=cs)
    sg = sampler.sample_gpu_fast(3)  # three isolated seeds: zero edges everywhere
    assert [l.e_size for l in sg.sampled_sgs] == [0, 0] and [l.src_size for l in sg.sampled_sgs] == [0, 0]
    assert np.array_equal(u32(sg.sampled_sgs[0].dev_column_offset), np.zeros(4, np.uint32))
    x = torch.zeros((0, 8), device="cuda")
    y = nts.SingleGPUAllSampleGraphOp(sg, 0, cs).forward(x)
    assert y.shape == (3, 8) and float(y.abs().sum()) == 0.0  # every output row is written (zeros)
    sg = sampler.sample_gpu_fast(1)  # seed 0: two in-neighbours, one of which has one
    assert sg.sampled_sgs[0].e_size == 2 and np.array_equal(u32(sg.sampled_sgs[0].dev_source), [1, 2])
    assert sg.sampled_sgs[1].e_size == 1 and np.array_equal(u32(sg.sampled_sgs[1].dev_source), [3])


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("F", [1, 7, 16, 41, 100, 128, 256, 602, 1433])
def test_gather_and_aggregate_match_oracle(nts, cs, F):
    V = 3000
    pairs, graph = make_graph(nts, cs, V, 25, seed=F)
    co, ri = oracle.build_csc(pairs, V)
    ind, outd = oracle.degrees(pairs, V)
    rng = np.random.default_rng(F)
    seeds = rng.permutation(V)[:400].astype(np.uint32)
    sampler = nts.FastSampler(graph, seeds, 2, 400, [8, 6], cuda_stream=cs)
    sg = sampler.sample_gpu_fast(400)
    lay = sg.sampled_sgs[1]
    r = dict(co=u32(lay.dev_column_offset), ri=u32(lay.dev_row_indices), ro=u32(lay.dev_row_offset), ci=u32(lay.dev_column_indices),
             wf=f32(lay.dev_edge_weight_forward), wb=f32(lay.dev_edge_weight_backward), src=u32(lay.dev_source))
    table_np = rng.standard_normal((V, F)).astype(np.float32)
    table = torch.from_numpy(table_np).cuda()
    x0 = torch.empty((lay.src_size, F), device="cuda")
    sampler.load_feature_gpu(cs, sg, x0, table)
    X0 = oracle.gather_rows(table_np, r["src"])
    assert np.array_equal(bits(f32(x0)), bits(X0))
    op = nts.SingleGPUAllSampleGraphOp(sg, 1, cs)
    y = op.forward(x0)
    Y = oracle.aggregate_fwd(X0, r["co"], r["ri"], r["wf"])
    np.testing.assert_allclose(f32(y), Y, rtol=RTOL, atol=1e-6)
    assert np.array_equal(bits(f32(y)), bits(Y))
    dY = rng.standard_normal(Y.shape).astype(np.float32)
    dy = torch.from_numpy(dY).cuda()
    dx = op.backward(dy)
    DX = oracle.aggregate_bwd_csr(dY, r["ro"], r["ci"], r["wb"])
    np.testing.assert_allclose(f32(dx), DX, rtol=RTOL, atol=1e-6)
    assert np.array_equal(bits(f32(dx)), bits(DX))
    dx2 = torch.full_like(dx, 7.0)  # push zeroes its output itself
    cs.Push_From_Dst_To_Src_Spmm(dy, dx2, lay.dev_e_w(), lay.dev_r_i(), lay.dev_c_o(), lay.src_size, 0, 0, 0, 0,
                                 lay.e_size, lay.v_size, F, True, False)
    np.testing.assert_allclose(f32(dx2), DX, rtol=1e-4, atol=1e-5)
    # unweighted (with_weight=false)
    y1 = torch.empty_like(y)
    cs.Gather_By_Dst_From_Src_Spmm(x0, y1, None, lay.dev_r_i(), lay.dev_c_o(), lay.src_size, 0, 0, 0, 0, lay.e_size,
                                   lay.v_size, F, False, False)
    np.testing.assert_allclose(f32(y1), oracle.aggregate_fwd(X0, r["co"], r["ri"], np.ones_like(r["wf"])), rtol=RTOL, atol=1e-6)
    # autograd glue
    xg = x0.clone().requires_grad_(True)
    (op(xg) * dy).sum().backward()
    np.testing.assert_allclose(f32(xg.grad), DX, rtol=RTOL, atol=1e-6)


@pytest.mark.parametrize("F,pitch", [(7, 7), (128, 128), (602, 608), (1433, 1433)])
def test_long_segments_take_the_block_path_with_identical_bits(nts, cs, F, pitch):
    """With nb_set_option("agg_long_rows", 1), hub rows (segments longer than 96 entries) are reduced by a whole block: same
    order, same bits as the oracle and as the default warp-per-row kernel, for every length around the threshold, with and
    without weights, padded or dense rows."""
    lib, check, ptr = nts._capi.lib(), nts._capi.check, nts._capi.ptr
    rng = np.random.default_rng(F)
    lens = np.array([0, 1, 95, 96, 97, 130, 0, 500, 3, 5000, 64, 257, 96, 2], np.uint32)
    off = np.zeros(lens.size + 1, np.uint32)
    off[1:] = np.cumsum(lens)
    E, R, S = int(off[-1]), lens.size, 900
    idx = rng.integers(0, S, E).astype(np.uint32)
    w = rng.standard_normal(E).astype(np.float32)
    X = rng.standard_normal((S, F)).astype(np.float32)
    xp = torch.zeros((S, pitch), device="cuda")
    xp[:, :F] = torch.from_numpy(X).cuda()
    d_off, d_idx, d_w = (torch.from_numpy(a.view(np.int32) if a.dtype == np.uint32 else a).cuda() for a in (off, idx, w))
    for weight, W in ((d_w, w), (None, np.ones(E, np.float32))):
        Y = oracle.aggregate_fwd(X, off, idx, W)
        outs = []
        for long_rows in (1, 0):
            check(lib.nb_set_option(b"agg_long_rows", long_rows))
            y = torch.full((R, pitch), 3.0, device="cuda")
            for _ in range(2):       # twice: the long-row list must re-arm itself between launches
                cs.aggregate_fwd_pitched(xp, y, weight, d_idx, d_off, R, F, pitch, pitch)
            outs.append(f32(y)[:, :F])
        check(lib.nb_set_option(b"agg_long_rows", 0))          # the default: plain warp-per-row kernel
        assert np.array_equal(bits(outs[0]), bits(Y)) and np.array_equal(bits(outs[1]), bits(Y))
    # the CSR backward entry point shares the kernels: rows = sources, entries = (dst, w_b)
    for long_rows in (1, 0):
        check(lib.nb_set_option(b"agg_long_rows", long_rows))
        dx = torch.empty((R, pitch), device="cuda")
        cs.aggregate_bwd_pitched(xp, dx, d_w, d_off, d_idx, R, F, pitch, pitch)
        assert np.array_equal(bits(f32(dx)[:, :F]), bits(oracle.aggregate_bwd_csr(X, off, idx, w)))


@pytest.mark.parametrize("F", [16, 100, 128, 256])
@pytest.mark.parametrize("R", [3, 1000, 6001])
def test_short_row_and_small_launch_kernels_keep_the_bits(nts, cs, F, R):
    """The CSR backward of rows <= 32 vectors wide runs 4 rows per warp (agg_short_rows), launches whose rows all fit on the
    GPU at once keep 16 / 8 entries in flight (agg_deep_small): both are scheduling only -- same bits as the oracle and as
    the plain warp-per-row kernel, for empty rows, 1-entry rows, rows around the joint-phase length and hub rows."""
    lib, check = nts._capi.lib(), nts._capi.check
    rng = np.random.default_rng(F * 7 + R)
    lens = rng.choice(np.array([0, 1, 1, 1, 2, 3, 4, 5, 9, 33, 70], np.uint32), R)
    lens[R // 2] = 300
    off = np.zeros(R + 1, np.uint32)
    off[1:] = np.cumsum(lens)
    E, S = int(off[-1]), 700
    idx = rng.integers(0, S, E).astype(np.uint32)
    w = rng.standard_normal(E).astype(np.float32)
    X = rng.standard_normal((S, F)).astype(np.float32)
    x = torch.from_numpy(X).cuda()
    d_off, d_idx, d_w = (torch.from_numpy(a.view(np.int32) if a.dtype == np.uint32 else a).cuda() for a in (off, idx, w))
    want = oracle.aggregate_fwd(X, off, idx, w)
    try:
        for short, deep in ((1, 1), (0, 0), (1, 0), (0, 1)):
            check(lib.nb_set_option(b"agg_short_rows", short))
            check(lib.nb_set_option(b"agg_deep_small", deep))
            y = torch.full((R, F), 3.0, device="cuda")
            cs.aggregate_fwd_pitched(x, y, d_w, d_idx, d_off, R, F, F, F)
            assert np.array_equal(bits(f32(y)), bits(want)), (short, deep, "fwd")
            dx = torch.full((R, F), 5.0, device="cuda")
            cs.aggregate_bwd_pitched(x, dx, d_w, d_off, d_idx, R, F, F, F)
            assert np.array_equal(bits(f32(dx)), bits(want)), (short, deep, "bwd")
            dx = torch.full((R, F), 5.0, device="cuda")
            cs.aggregate_bwd_pitched(x, dx, None, d_off, d_idx, R, F, F, F)
            assert np.array_equal(bits(f32(dx)), bits(oracle.aggregate_fwd(X, off, idx, np.ones(E, np.float32)))), (short, deep, "bwd, no weights")
    finally:
        check(lib.nb_set_option(b"agg_short_rows", 0))      # the defaults
        check(lib.nb_set_option(b"agg_deep_small", 1))


@pytest.mark.parametrize("V,E,seed", [(1, 5, 0), (7, 0, 1), (2708, 13566, 2), (300, 70000, 3), (70000, 300000, 4), (20_000_000, 3_000_000, 5),
                                      (232965, 12_000_000, 6)])
def test_graph_built_on_device_matches_reference_ordering(nts, cs, V, E, seed):
    """nb_graph_create_from_pairs (stable radix sort by dst on the device) == FullyRepGraph::GenerateAll's ordering as restated
    by the oracle: column = dst, entries in FILE order, degrees clamped >= 1. Multi-edges, hubs, empty columns, 1..4 digit passes."""
    rng = np.random.default_rng(seed)
    src = (rng.random(E) ** 3 * V).astype(np.uint32)            # skewed: hubs and many duplicates
    dst = (rng.random(E) ** 2 * V).astype(np.uint32)
    pairs = np.stack([src, dst], 1).astype(np.uint32)
    co, ri = oracle.build_csc(pairs, V)
    ind, outd = oracle.degrees(pairs, V)
    for on_device in (False, True):
        ep = torch.from_numpy(pairs.view(np.int32)).cuda() if on_device else pairs
        g = nts.FullyRepGraph(cs, V, edge_pairs=ep)
        d_co, d_ri, d_in, d_out = g.device_arrays()
        assert np.array_equal(u32(d_co), co) and (E == 0 or np.array_equal(u32(d_ri)[:E], ri))
        assert np.array_equal(u32(d_in), ind) and np.array_equal(u32(d_out), outd)
    if E:
        h = nts.FullyRepGraph(cs, V, edge_pairs=pairs, build_on_host=True)
        assert np.array_equal(u32(h.device_arrays()[1])[:E], ri)
    if V > 1 and E:
        bad = pairs.copy()
        bad[E // 2, 0] = V
        with pytest.raises(nts.NtsError):
            nts.FullyRepGraph(cs, V, edge_pairs=bad)


def test_unaligned_views_fall_back_to_narrower_vectors(nts, cs):
    V, F = 500, 64
    pairs, graph = make_graph(nts, cs, V, 10, seed=1)
    rng = np.random.default_rng(0)
    big = torch.from_numpy(rng.standard_normal(V * F + 3).astype(np.float32)).cuda()
    ids = torch.from_numpy(rng.integers(0, V, 300).astype(np.int32)).cuda()
    for shift in (0, 1, 2):
        table = big[shift:shift + V * F].view(V, F)
        out = torch.empty((300, F), device="cuda")
        cs.zero_copy_feature_move_gpu(out, table, ids, F, 300)
        assert torch.equal(out, table[ids.long()])


def test_cached_gather_labels_and_row_override(nts, cs):
    V, F, E2 = 6000, 100, 48
    rng = np.random.default_rng(4)
    table_np = rng.standard_normal((V, F)).astype(np.float32)
    hot = np.sort(rng.permutation(V)[:1500]).astype(np.uint32)
    cache_np = table_np[hot] + np.float32(1000.0)  # deliberately different so the source of each row is visible
    hashmap = np.full(V, 0xFFFFFFFF, np.uint32)
    hashmap[hot] = np.arange(hot.size, dtype=np.uint32)
    ids_np = rng.integers(0, V, 5000).astype(np.uint32)
    ref = oracle.gather_rows_cached(table_np, cache_np, hashmap, ids_np)
    # cold table in mapped pinned host memory, as the reference keeps it (core/ntsDataloador.hpp:187)
    cold = torch.from_numpy(table_np).pin_memory()
    out = torch.empty((5000, F), device="cuda")
    hits = torch.zeros(1, dtype=torch.int32, device="cuda")
    cs.gather_feature_cached(out, cold, torch.from_numpy(cache_np).cuda(), torch.from_numpy(hashmap.view(np.int32)).cuda(),
                             torch.from_numpy(ids_np.view(np.int32)).cuda(), F, 5000, hits)
    cs.CUDA_DEVICE_SYNCHRONIZE()
    assert np.array_equal(bits(f32(out)), bits(ref))
    assert int(hits.item()) == int((hashmap[ids_np] != 0xFFFFFFFF).sum())
    labels = rng.integers(0, 41, V).astype(np.int64)
    lab = torch.empty(5000, dtype=torch.int64, device="cuda")
    cs.global_copy_label_move_gpu(lab, torch.from_numpy(labels).cuda(), torch.from_numpy(ids_np.view(np.int32)).cuda(), 5000)
    assert np.array_equal(lab.cpu().numpy(), oracle.gather_labels(labels, ids_np))
    # hot-row override for super batch 2
    cache_map = np.full(V, 0xFFFFFFFF, np.uint32)
    cache_loc = np.zeros(V, np.uint32)
    oracle.set_cache_index(cache_map, cache_loc, 2, hot)
    other = rng.permutation(V)[:300].astype(np.uint32)
    cache_map[other] = 1
    dst = rng.permutation(V)[:2000].astype(np.uint32)
    emb = rng.standard_normal((2000, E2)).astype(np.float32)
    feat = rng.standard_normal((2000, F)).astype(np.float32)
    share_e = rng.standard_normal((hot.size, E2)).astype(np.float32)
    share_f = rng.standard_normal((hot.size, F)).astype(np.float32)
    d = lambda a: torch.from_numpy(a.view(np.int32) if a.dtype == np.uint32 else a).cuda()
    te, tf_ = d(emb.copy()), d(feat.copy())
    cs.dev_load_share_embedding(te, d(share_e), d(cache_map), d(cache_loc), E2, d(dst), 2000, 2)
    assert np.array_equal(f32(te), oracle.row_override(emb, share_e, cache_map, cache_loc, dst, 2))
    te = d(emb.copy())
    cs.dev_load_share_embedding_and_feature(tf_, te, d(share_f), d(share_e), d(cache_map), d(cache_loc), F, E2, d(dst), 2000, 2)
    assert np.array_equal(f32(te), oracle.row_override(emb, share_e, cache_map, cache_loc, dst, 2))
    assert np.array_equal(f32(tf_), oracle.row_override(feat, share_f, cache_map, cache_loc, dst, 2))


@pytest.mark.parametrize("F,hot_frac", [(100, 0.25), (602, 0.1), (1433, 0.2), (7, 0.0), (128, 1.0)])
def test_cached_load_pair_in_the_reference_call_shape(nts, cs, F, hot_frac):
    """FastSampler::load_feature_gpu_cache as the reference issues it (core/ntsFastSampler.hpp:284-312): a CPU split of the bottom
    layer's sources into a cold and a hot position list held in MAPPED PINNED host arrays, then one indexed gather per list. The two
    calls together must equal the oracle's cached gather, bit for bit (and the fused nb_gather_rows_cached)."""
    V, S = 5000, 3000
    rng = np.random.default_rng(11)
    table_np = rng.standard_normal((V, F)).astype(np.float32)
    hot = np.sort(rng.permutation(V)[:int(V * hot_frac)]).astype(np.uint32)
    cache_np = table_np[hot] + np.float32(1000.0) if hot.size else np.zeros((1, F), np.float32)
    hashmap = np.full(V, 0xFFFFFFFF, np.uint32)
    hashmap[hot] = np.arange(hot.size, dtype=np.uint32)
    src_np = np.sort(rng.permutation(V)[:S]).astype(np.uint32)          # dev_source: distinct, ascending
    ref = oracle.gather_rows_cached(table_np, cache_np, hashmap, src_np)
    is_hot = hashmap[src_np] != 0xFFFFFFFF
    pin = lambda a: torch.from_numpy(a.view(np.int32) if a.dtype == np.uint32 else a).pin_memory()
    local_idx, local_idx_cache = pin(np.flatnonzero(~is_hot).astype(np.uint32)), pin(np.flatnonzero(is_hot).astype(np.uint32))
    h_hashmap, cold = pin(hashmap), pin(table_np)
    out = torch.full((S, F), float("nan"), device="cuda")
    src = torch.from_numpy(src_np.view(np.int32)).cuda()
    cs.zero_copy_feature_move_gpu_cache(out, cold, src, F, int(local_idx.numel()), local_idx)
    cs.gather_feature_from_gpu_cache(out, torch.from_numpy(cache_np).cuda(), src, F, int(local_idx_cache.numel()), local_idx_cache, h_hashmap)
    assert np.array_equal(bits(f32(out)), bits(ref))


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("F", [8, 41, 128])
def test_gat_legacy_ops_and_fused_layer(nts, cs, F):
    V = 5000
    pairs, graph = make_graph(nts, cs, V, 30, seed=F + 1)
    rng = np.random.default_rng(F)
    seeds = rng.permutation(V)[:500].astype(np.uint32)
    sampler = nts.FastSampler(graph, seeds, 2, 500, [12, 40], cuda_stream=cs, merge_src_dst=True, build_csr=True)
    sg = sampler.sample_gpu_fast(500, weightType=nts.WeightType.None_)
    for hop in (0, 1):
        lay = sg.sampled_sgs[hop]
        co, ri, dl = u32(lay.dev_column_offset), u32(lay.dev_row_indices), u32(lay.dev_dst_local_id)
        H = rng.standard_normal((lay.src_size, F)).astype(np.float32)
        att = (rng.standard_normal(2 * F) * 0.3).astype(np.float32)
        h, a = torch.from_numpy(H).cuda(), torch.from_numpy(att).cuda()
        # legacy chain
        msg = nts.BatchGPUSrcDstScatterOp(sg, hop, cs).forward(h)
        MSG = oracle.scatter_src_dst(H, co, ri, dl)
        assert np.array_equal(f32(msg), MSG)
        m_np = MSG @ att
        m_np = np.where(m_np > 0, m_np, np.float32(0.2) * m_np).astype(np.float32)
        sm = nts.BatchGPUEdgeSoftMax(sg, hop, cs)
        alpha_t = sm.forward(torch.from_numpy(m_np).cuda().view(-1, 1))
        A = oracle.edge_softmax_fwd(m_np, co)
        np.testing.assert_allclose(f32(alpha_t).ravel(), A, rtol=RTOL, atol=1e-7)
        da = rng.standard_normal(A.shape).astype(np.float32)
        dm = sm.backward(torch.from_numpy(da).cuda().view(-1, 1))
        np.testing.assert_allclose(f32(dm).ravel(), oracle.edge_softmax_bwd(da, f32(alpha_t).ravel(), co), rtol=1e-4, atol=1e-6)
        agg = nts.BatchGPUAggregateDst(sg, hop, cs)
        emo = MSG[:, :F] * A[:, None]
        nbr = agg.forward(torch.from_numpy(emo).cuda())
        np.testing.assert_allclose(f32(nbr), oracle.gather_msg_to_dst(emo, co), rtol=RTOL, atol=1e-5)
        dn = rng.standard_normal((lay.v_size, F)).astype(np.float32)
        assert np.array_equal(f32(agg.backward(torch.from_numpy(dn).cuda())), oracle.scatter_dst_to_msg(dn, co))
        dmsg = rng.standard_normal(MSG.shape).astype(np.float32)
        gx = nts.BatchGPUSrcDstScatterOp(sg, hop, cs).backward(torch.from_numpy(dmsg).cuda())
        np.testing.assert_allclose(f32(gx), oracle.gather_src_dst(dmsg, co, ri, dl, lay.src_size), rtol=1e-4, atol=1e-4)
        # fused layer vs the oracle's restatement of the toolkit's chain
        op = nts.GATFusedOp(sg, hop, cs)
        out = op.forward(h, a)
        OUT, ALPHA, PRE = oracle.gat_layer_fwd(H, att, co, ri, dl)
        np.testing.assert_allclose(f32(out), OUT, rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(f32(op.alpha), ALPHA, rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(f32(op.score_pre), PRE, rtol=1e-4, atol=1e-5)
        dout = rng.standard_normal(OUT.shape).astype(np.float32)
        dh, datt = op.backward(h, a, torch.from_numpy(dout).cuda())
        DH, DATT = oracle.gat_layer_bwd(H, att, dout, f32(op.score_pre), f32(op.alpha), co, ri, dl)
        np.testing.assert_allclose(f32(dh), DH, rtol=1e-3, atol=1e-4)
        np.testing.assert_allclose(f32(datt), DATT, rtol=1e-3, atol=1e-3)
        # autograd glue
        hg, ag = h.clone().requires_grad_(True), a.clone().requires_grad_(True)
        (op(hg, ag) * torch.from_numpy(dout).cuda()).sum().backward()
        np.testing.assert_allclose(f32(hg.grad), DH, rtol=1e-3, atol=1e-4)


def _gat_goldens():
    import glob
    import os
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    return sorted(os.path.basename(f)[4:-4] for f in glob.glob(os.path.join(d, "gat_*.npz")))


@pytest.fixture(params=[0, 1])
def gat_short_rows(nts, request):
    """the fused GAT backward's CSR reduction through the warp-per-row kernel (default) and the 4-rows-per-warp one"""
    nts._capi.check(nts._capi.lib().nb_set_option(b"agg_short_rows", request.param))
    yield request.param
    nts._capi.check(nts._capi.lib().nb_set_option(b"agg_short_rows", 0))


@pytest.mark.parametrize("name", _gat_goldens())
def test_gat_against_reference_kernel_records(nts, cs, name, gat_short_rows):
    """a14 pinned: tests/golden/gat_*.npz hold the outputs of the reference's OWN CUDA kernels (cuda/ntsCUDADistKernel.cuh, run on a
    B200 by oracle/_ref/ref_gpu_driver <- oracle/make_gat_golden.py) for the op chain of toolkits/GAT_SAMPLE_ALL_MULTI.hpp:383-464.
    Legacy-shaped ops and the fused layer, forward and backward, against those records. Tolerance: 1e-5 relative (north star) plus an
    absolute floor of 1e-5 x the tensor's largest magnitude -- the reference sums in a different order (cuBLAS dot over [E,2F] for the
    score, float atomics for the source gradient), so elements that cancel to near zero cannot agree to 1e-5 of themselves."""
    import os
    lib, check, ptr = nts._capi.lib(), nts._capi.check, nts._capi.ptr
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"gat_{name}.npz")))
    F, S = int(g["F"]), int(g["n_src"])
    co, ri, dl = g["column_offset"], g["row_indices"], g["dst_local_id"]
    V, E = co.size - 1, ri.size
    d32 = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int32) if a.dtype == np.uint32 else np.ascontiguousarray(a)).cuda()
    t_co, t_ri, t_dl = d32(co), d32(ri), d32(dl)
    h, att, dout = d32(g["h"].reshape(S, F)), d32(g["att"]), d32(g["dout"].reshape(V, F))

    def close(mine, ref, what, rtol=1e-5, floor=1e-5):
        ref = np.asarray(ref, np.float32).reshape(mine.shape)
        np.testing.assert_allclose(mine, ref, rtol=rtol, atol=floor * float(np.abs(ref).max() + 1e-30), err_msg=f"{name}: {what}")

    # ---- legacy-shaped ops, one reference kernel each
    alpha = torch.empty(E, device="cuda"); cached = torch.empty(E, device="cuda")
    t_m, t_da, t_alpha_ref = d32(g["m"]), d32(g["d_a"]), d32(g["alpha"])
    cs.Edge_Softmax_Forward_Norm_Block(alpha, t_m, cached, t_ri, t_co, V, 1)
    close(f32(alpha), g["alpha"], "Edge_Softmax_Forward_Norm_Block")
    assert torch.equal(alpha, cached)
    d_m = torch.empty(E, device="cuda")
    cs.Edge_Softmax_Backward_Block(d_m, t_da, t_alpha_ref, t_ri, t_co, V, 1)
    close(f32(d_m), g["d_m"], "Edge_Softmax_Backward_Block")
    e_src = g["h"].reshape(S, F)[ri.astype(np.int64)]
    emo = (e_src * g["alpha"][:, None]).astype(np.float32)               # e_msg.slice(0:F) * a, one rounding per element
    nbr = torch.empty((V, F), device="cuda")
    t_emo = d32(emo)
    cs.Gather_Msg_to_Dst(nbr, t_emo, t_ri, t_co, V, F)
    close(f32(nbr), g["out"], "Gather_Msg_to_Dst")
    if "e_msg" in g:                                                     # small cases carry the per-edge tensors
        msg = torch.empty((E, 2 * F), device="cuda")
        cs.Scatter_Src_Dst_to_Msg(msg, h, t_ri, t_co, V, F, t_dl)
        assert np.array_equal(f32(msg), g["e_msg"].reshape(E, 2 * F)), "Scatter_Src_Dst_to_Msg is a copy: bit-exact"
        assert np.array_equal(emo, g["e_msg_out"].reshape(E, F))
        back = torch.empty((E, F), device="cuda")
        cs.Scatter_Dst_to_Msg(back, dout, t_ri, t_co, V, F)
        assert np.array_equal(f32(back), g["d_e_msg_out"].reshape(E, F)), "Scatter_Dst_to_Msg is a copy: bit-exact"
        gsrc = torch.zeros((S, F), device="cuda")
        t_dmsg = d32(g["d_e_msg"].reshape(E, 2 * F))
        cs.Gather_Msg_To_Src_Dst(gsrc, t_dmsg, t_ri, t_co, V, F, t_dl, S)
        close(f32(gsrc), g["dh"], "Gather_Msg_To_Src_Dst")
    # ---- fused layer (nb_gat_fwd / nb_gat_bwd) against the same records
    pre, al, out = torch.empty(E, device="cuda"), torch.empty(E, device="cuda"), torch.empty((V, F), device="cuda")
    check(lib.nb_gat_fwd(cs._h, ptr(h), ptr(att), 0.2, ptr(t_co), ptr(t_ri), ptr(t_dl), V, S, F, ptr(pre), ptr(al), ptr(out)))
    close(f32(pre), g["score_pre"], "fused score")
    close(f32(al), g["alpha"], "fused alpha")
    close(f32(out), g["out"], "fused forward output")
    order = np.argsort(ri, kind="stable").astype(np.uint32)              # stable CSR of the layer (sampCSC::csc_to_csr order)
    row_offset = np.zeros(S + 1, np.uint32)
    np.cumsum(np.bincount(ri, minlength=S), out=row_offset[1:])
    edge_dst = np.repeat(np.arange(V, dtype=np.uint32), np.diff(co.astype(np.int64)))
    src_to_dst = np.full(S, 0xFFFFFFFF, np.uint32)
    src_to_dst[dl] = np.arange(V, dtype=np.uint32)
    dh, datt = torch.empty((S, F), device="cuda"), torch.empty(2 * F, device="cuda")
    t_ro, t_ci, t_c2c, t_s2d = d32(row_offset), d32(edge_dst[order]), d32(order), d32(src_to_dst)   # named: they must outlive the calls
    args = (cs._h, ptr(h), ptr(att), 0.2, ptr(dout), ptr(pre), ptr(al), ptr(t_co), ptr(t_ri), ptr(t_dl), ptr(t_ro),
            ptr(t_ci), ptr(t_c2c), ptr(t_s2d), V, S, E, F)
    check(lib.nb_gat_bwd(*args, ptr(dh), ptr(datt)))
    close(f32(dh), g["dh"], "fused dH")
    close(f32(datt), g["datt"], "fused d(att)", floor=2e-5)
    dh2, datt2 = torch.empty_like(dh), torch.empty_like(datt)
    check(lib.nb_gat_bwd(*args, ptr(dh2), ptr(datt2)))                   # no float atomics anywhere: run-to-run identical
    assert torch.equal(dh, dh2) and torch.equal(datt, datt2)


def test_gat_fused_layer_with_hub_sources(nts, cs):
    """Skewed sources: CSR rows of hundreds of entries (block-per-row segment reduction, warp-cooperative row pass in the backward)."""
    V, F = 600, 128
    rng = np.random.default_rng(77)
    p = 1.0 / np.arange(1, V + 1) ** 1.2
    p /= p.sum()
    pairs = np.concatenate([np.stack([rng.choice(V, 60, replace=False, p=p), np.full(60, v)], 1) for v in range(V)]).astype(np.uint32)
    graph = nts.FullyRepGraph(cs, V, edge_pairs=pairs)
    seeds = rng.permutation(V)[:400].astype(np.uint32)
    sampler = nts.FastSampler(graph, seeds, 2, 400, [12, 40], cuda_stream=cs, merge_src_dst=True, build_csr=True)
    sg = sampler.sample_gpu_fast(400, weightType=nts.WeightType.None_)
    for hop, long_rows in ((0, 0), (1, 0), (1, 1)):
        nts._capi.check(nts._capi.lib().nb_set_option(b"agg_long_rows", long_rows))
        lay = sg.sampled_sgs[hop]
        co, ri, dl = u32(lay.dev_column_offset), u32(lay.dev_row_indices), u32(lay.dev_dst_local_id)
        assert np.diff(u32(lay.dev_row_offset).astype(np.int64)).max() > (96 if hop == 1 else 32)
        H = rng.standard_normal((lay.src_size, F)).astype(np.float32)
        att = (rng.standard_normal(2 * F) * 0.3).astype(np.float32)
        h, a = torch.from_numpy(H).cuda(), torch.from_numpy(att).cuda()
        op = nts.GATFusedOp(sg, hop, cs)
        out = op.forward(h, a)
        OUT, ALPHA, PRE = oracle.gat_layer_fwd(H, att, co, ri, dl)
        np.testing.assert_allclose(f32(out), OUT, rtol=1e-4, atol=1e-5)
        dout = rng.standard_normal(OUT.shape).astype(np.float32)
        dh, datt = op.backward(h, a, torch.from_numpy(dout).cuda())
        DH, DATT = oracle.gat_layer_bwd(H, att, dout, f32(op.score_pre), f32(op.alpha), co, ri, dl)
        np.testing.assert_allclose(f32(dh), DH, rtol=1e-3, atol=2e-4)
        np.testing.assert_allclose(f32(datt), DATT, rtol=1e-3, atol=2e-3)
    nts._capi.check(nts._capi.lib().nb_set_option(b"agg_long_rows", 0))


# ---------------------------------------------------------------------------------------------
def reddit_shaped(V=232965, E=114615892, seed=0x5EED0001):
    """BASELINE.json configs[1] shape, generated on the GPU: power-law in-degree (mean ~492), skewed sources."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    w = torch.rand(V, generator=g, device="cuda").clamp_min(1e-6).pow(-0.65)
    deg = (w / w.sum() * E).floor().clamp_(1, V - 1).to(torch.int64)
    deg[0] += E - int(deg.sum()) if int(deg.sum()) < E else 0
    co = torch.zeros(V + 1, dtype=torch.int64, device="cuda")
    co[1:] = deg.cumsum(0)
    total = int(co[-1])
    src = (torch.rand(total, generator=g, device="cuda").pow(1.6) * V).to(torch.int64).clamp_(0, V - 1)
    return co.to(torch.int32).cpu().numpy().view(np.uint32), src.to(torch.int32).cpu().numpy().view(np.uint32)


def test_full_size_reddit_shaped_properties(nts, cs):
    """At BASELINE.json's size (233K vertices, ~114M edges, F=602, batch 1024, fanout 25-10): properties
    that do not need the oracle to finish -- structural invariants and linearity / checksum identities."""
    V, F = 232965, 602
    co, ri = reddit_shaped()
    graph = nts.FullyRepGraph(cs, V, column_offset=co, row_indices=ri)
    rng = np.random.default_rng(3)
    seeds = rng.permutation(V)[:1024].astype(np.uint32)
    sampler = nts.FastSampler(graph, seeds, 2, 1024, [25, 10], cuda_stream=cs)
    sg = sampler.sample_gpu_fast(1024)
    top, bottom = sg.sampled_sgs
    assert top.v_size == 1024 and bottom.v_size == top.src_size
    assert torch.equal(bottom.dev_destination, top.dev_source)  # layer chaining invariant
    deg = np.diff(co.astype(np.int64))
    for lay, f in ((top, 25), (bottom, 10)):
        lens = np.diff(u32(lay.dev_column_offset).astype(np.int64))
        assert np.array_equal(lens, np.minimum(deg[u32(lay.dev_destination)], f))
        src = u32(lay.dev_source).astype(np.int64)
        assert (np.diff(src) > 0).all()  # ascending, unique
        assert np.array_equal(np.unique(u32(lay.dev_sample_ans)), src)  # exactly the sampled set
        assert np.array_equal(src[u32(lay.dev_row_indices)], u32(lay.dev_sample_ans))
        ro, ci, c2c = u32(lay.dev_row_offset), u32(lay.dev_column_indices), u32(lay.dev_csr_to_csc)
        assert ro[-1] == lay.e_size and np.array_equal(np.sort(c2c), np.arange(lay.e_size, dtype=np.uint32))
        assert np.array_equal(u32(lay.dev_row_indices)[c2c], np.repeat(np.arange(lay.src_size), np.diff(ro.astype(np.int64))))
        e_dst = np.repeat(np.arange(lay.v_size), lens)
        assert np.array_equal(ci, e_dst[c2c])
        for s in np.nonzero(np.diff(ro.astype(np.int64)) > 1)[0][:2000]:
            assert (np.diff(c2c[ro[s]:ro[s + 1]].astype(np.int64)) > 0).all()  # stable order inside a row
        assert np.array_equal(f32(lay.dev_edge_weight_backward), f32(lay.dev_edge_weight_forward)[c2c])
    # gather + aggregate identities on the bottom layer at F=602
    table = torch.randn((V, F), device="cuda")
    x0 = torch.empty((bottom.src_size, F), device="cuda")
    sampler.load_feature_gpu(cs, sg, x0, table)
    assert torch.equal(x0, table[bottom.dev_source.long()])
    op = nts.SingleGPUAllSampleGraphOp(sg, 1, cs)
    y = op.forward(x0)
    # checksum of checksums against an independent fp64 formulation
    w = bottom.dev_edge_weight_forward.double()
    e_dst = torch.repeat_interleave(torch.arange(bottom.v_size, device="cuda"), torch.from_numpy(np.diff(u32(bottom.dev_column_offset).astype(np.int64))).cuda())
    ref = torch.zeros((bottom.v_size, F), dtype=torch.float64, device="cuda")
    ref.index_add_(0, e_dst, x0.double()[bottom.dev_row_indices.long()] * w[:, None])
    torch.testing.assert_close(y.double(), ref, rtol=1e-5, atol=1e-5)
    # linearity: A(2x + z) = 2A(x) + A(z)
    z = torch.randn_like(x0)
    torch.testing.assert_close(op.forward(2 * x0 + z), 2 * y + op.forward(z), rtol=1e-4, atol=1e-4)
    # adjointness: <A x, g> = <x, A^T g>
    gy = torch.randn_like(y)
    gx = op.backward(gy)
    lhs, rhs = (y.double() * gy.double()).sum(), (x0.double() * gx.double()).sum()
    assert abs(float(lhs - rhs)) <= 1e-6 * float((y.double().abs() * gy.double().abs()).sum())
    refb = torch.zeros((bottom.src_size, F), dtype=torch.float64, device="cuda")
    refb.index_add_(0, bottom.dev_row_indices.long(), gy.double()[e_dst] * w[:, None])
    torch.testing.assert_close(gx.double(), refb, rtol=1e-5, atol=1e-5)


# ---------------------------------------------------------------------------------------------
def test_stage_by_stage_entry_points_match_oracle(nts, cs):
    """The reference's own call sequence (SampledSubgraph::gpu_sampling_init_co -> gpu_sampling -> update_degrees_GPU ->
    Get_Weight, core/FullyRepGraph.hpp:213-239, 326-524) through the stage-shaped entry points."""
    import ctypes as C
    lib, check, ptr = nts._capi.lib(), nts._capi.check, nts._capi.ptr
    V = 8000
    pairs, graph = make_graph(nts, cs, V, 25, seed=21)
    co, ri = oracle.build_csc(pairs, V)
    ind, outd = oracle.degrees(pairs, V)
    g_co, g_ri, g_in, g_out = graph.device_arrays()
    rng = np.random.default_rng(9)
    dst_np = rng.permutation(V)[:600].astype(np.uint32)
    dev = lambda a: torch.from_numpy(a.view(np.int32) if a.dtype == np.uint32 else a).cuda()
    for fanout, merge in ((7, False), (40, True)):
        dst = dev(dst_np)
        lco = torch.zeros(dst_np.size + 1, dtype=torch.int32, device="cuda")
        e = C.c_uint32()
        check(lib.nb_sample_count(cs._h, ptr(dst), ptr(lco), ptr(g_co), dst_np.size, fanout, None, 0, C.byref(e)))
        ref_co, ref_e = oracle.count_offsets(dst_np, co, fanout)
        assert e.value == ref_e and np.array_equal(u32(lco), ref_co)
        r_i = torch.zeros(max(e.value, 1), dtype=torch.int32, device="cuda")
        src_index = torch.full((V,), -1, dtype=torch.int32, device="cuda")
        src = torch.zeros(e.value + dst_np.size, dtype=torch.int32, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        check(lib.nb_sample_traverse(cs._h, ptr(dst), ptr(lco), ptr(r_i), ptr(g_co), ptr(g_ri), ptr(src_index), dst_np.size,
                                     e.value, V, ptr(src), ptr(cnt), 0, fanout, 1 if merge else 0, 77, 5))
        ans = u32(r_i)[:e.value].copy()
        S = int(cnt.item())
        rsrc, rri, rdl = oracle.reindex(ans, dst_np, V, merge_src_dst=merge)
        assert S == rsrc.size and np.array_equal(u32(src)[:S], rsrc)
        assert np.array_equal(u32(src_index)[rsrc], np.arange(S, dtype=np.uint32))
        check(lib.nb_sample_update_ri(cs._h, ptr(r_i), ptr(src_index), e.value))
        assert np.array_equal(u32(r_i)[:e.value], rri)
        if merge:
            dl = torch.zeros(dst_np.size, dtype=torch.int32, device="cuda")
            check(lib.nb_set_dst_local_index(cs._h, ptr(src_index), ptr(dst), dst_np.size, ptr(dl)))
            assert np.array_equal(u32(dl), rdl)
        # sampling rule
        for j, d in enumerate(dst_np[:200]):
            nb = ri[co[d]:co[d + 1]]
            got = ans[ref_co[j]:ref_co[j + 1]]
            assert np.array_equal(got, nb) if nb.size <= fanout else (np.unique(got).size == fanout and np.isin(got, nb).all())
        # weights from the graph's degrees, then from per-batch sampled degrees
        w = torch.zeros(max(e.value, 1), device="cuda")
        check(lib.nb_edge_weight(cs._h, ptr(w), ptr(g_out), ptr(g_in), dst_np.size, ptr(dst), ptr(src), ptr(lco), ptr(r_i), 0))
        ewf, _ = oracle.weights(dst_np, rsrc, ref_co, rri, None, None, ind, outd, 0)
        assert np.array_equal(bits(f32(w)[:e.value]), bits(ewf))
        d_in, d_out = torch.full((V,), 9, dtype=torch.int32, device="cuda"), torch.full((V,), 9, dtype=torch.int32, device="cuda")
        check(lib.nb_update_degree(cs._h, ptr(d_out), ptr(d_in), V, dst_np.size, ptr(dst), ptr(src), ptr(lco), ptr(r_i), 0))
        s_in, s_out = oracle.update_degrees(dst_np, rsrc, ref_co, rri, V)
        assert np.array_equal(u32(d_in), s_in) and np.array_equal(u32(d_out), s_out)
        check(lib.nb_edge_weight(cs._h, ptr(w), ptr(d_out), ptr(d_in), dst_np.size, ptr(dst), ptr(src), ptr(lco), ptr(r_i), 1))
        ewf2, _ = oracle.weights(dst_np, rsrc, ref_co, rri, None, None, s_in, s_out, 2)
        assert np.array_equal(bits(f32(w)[:e.value]), bits(ewf2))


def test_padded_rows_and_tma_gather_match_dense(nts, cs):
    """Row-padded layouts (pitch 608 for 602-wide rows) take the 128-bit / TMA bulk-copy paths and must give the same bits."""
    lib, check, ptr = nts._capi.lib(), nts._capi.check, nts._capi.ptr
    V, F, P = 5000, 602, 608
    pairs, graph = make_graph(nts, cs, V, 25, seed=5)
    rng = np.random.default_rng(1)
    seeds = rng.permutation(V)[:300].astype(np.uint32)
    sampler = nts.FastSampler(graph, seeds, 2, 300, [8, 6], cuda_stream=cs)
    sg = sampler.sample_gpu_fast(300)
    lay = sg.sampled_sgs[1]
    dense = torch.randn((V, F), device="cuda")
    padded = torch.zeros((V, P), device="cuda")
    padded[:, :F] = dense
    x_d = torch.empty((lay.src_size, F), device="cuda")
    sampler.load_feature_gpu(cs, sg, x_d, dense)
    for variant in (0, 1):
        check(lib.nb_set_option(b"gather_variant", variant))
        x_p = torch.full((lay.src_size, P), 7.0, device="cuda")
        sampler.load_feature_gpu(cs, sg, x_p[:, :F], padded[:, :F])
        assert torch.equal(x_p[:, :F], x_d)
    op = nts.SingleGPUAllSampleGraphOp(sg, 1, cs)
    y_d = op.forward(x_d)
    y_p = op.forward(x_p[:, :F])
    assert y_p.stride(0) == P and torch.equal(y_p, y_d)
    dy = torch.randn((lay.v_size, P), device="cuda")
    g_p = op.backward(dy[:, :F])
    g_d = op.backward(dy[:, :F].contiguous())
    assert torch.equal(g_p, g_d)
    # the bottom hop straight from the feature table (no X0): identical bits
    y_f = torch.empty((lay.v_size, P), device="cuda")
    cs.aggregate_fwd_pitched(padded, y_f, lay.dev_e_w(), lay.dev_sample_ans, lay.dev_c_o(), lay.v_size, F, P, P)
    assert torch.equal(y_f[:, :F], y_d)
    # the same through the operator surface: load_feature_gpu(lazy=True) hands the op a promise instead of X0
    for tab in (dense, padded[:, :F]):
        buf = torch.full((lay.src_size + 3, tab.stride(0)), 5.0, device="cuda")[:, :F]
        lz = sampler.load_feature_gpu(cs, sg, buf, tab, lazy=True)
        assert isinstance(lz, nts.LazyFeature) and lz.shape == (lay.src_size, F)
        y_l = op.forward(lz)
        assert torch.equal(y_l, y_d) and bool((buf == 5.0).all())          # nothing was copied
        assert torch.equal(op(lz), y_d)                                     # autograd wrapper accepts the promise too
        assert torch.equal(lz.materialize(), x_d)                           # any other reader gets the ordinary gather
        top = nts.SingleGPUAllSampleGraphOp(sg, 0, cs)                      # a promise handed to the wrong hop is materialised first
        assert top.forward(torch.ones((sg.sampled_sgs[0].src_size, 4), device="cuda")).shape[0] == sg.sampled_sgs[0].v_size


@pytest.mark.parametrize("F,pitch", [(128, 128), (100, 100), (64, 64), (256, 256), (602, 608), (608, 608), (512, 520), (1000, 1000)])
@pytest.mark.parametrize("n", [1, 3, 4, 1001, 20000])
def test_tensor_map_gather4_variant_copies_the_same_rows(nts, cs, F, pitch, n):
    """nb_set_option("gather_variant", 2): rows move through TMA tensor maps, four per instruction (tile::gather4), and leave through
    tiled tensor stores. Every eligible shape gives the bits of a plain index; groups that are not a multiple of four rows are clipped
    at the output map's edge (the rows behind them are not touched); shapes it cannot take fall back."""
    lib, check, ptr = nts._capi.lib(), nts._capi.check, nts._capi.ptr
    V = 50000
    g = torch.Generator(device="cuda").manual_seed(F * 131 + n)
    table = torch.zeros((V, pitch), device="cuda")
    table[:, :F] = torch.randn((V, F), device="cuda", generator=g)
    ids = torch.randint(0, V, (n,), device="cuda", dtype=torch.int32, generator=g)
    out = torch.full((n + 5, pitch), 9.0, device="cuda")
    try:
        check(lib.nb_set_option(b"gather_variant", 2))
        check(lib.nb_gather_rows(cs._h, ptr(out), ptr(table), ptr(ids), n, F, pitch, pitch))
        cs.CUDA_DEVICE_SYNCHRONIZE()
    finally:
        check(lib.nb_set_option(b"gather_variant", 1))
    assert torch.equal(out[:n, :F], table[ids.long()][:, :F])
    assert bool((out[n:] == 9.0).all())


def test_async_sampling_pipeline_slots_and_no_bottom_csr(nts, cs):
    """sample_gpu_fast(sync=False) + wait(): two pipeline slots in flight give the same subgraphs as the synchronous call;
    bottom_csr=False drops only the bottom layer's CSR."""
    V = 6000
    pairs, graph = make_graph(nts, cs, V, 30, seed=17)
    seeds = np.random.default_rng(3).permutation(V)[:1024].astype(np.uint32)
    a = nts.FastSampler(graph, seeds, 2, 256, [9, 5], pipeline_num=2, cuda_stream=[cs, cs], rng_seed=5, bottom_csr=False)
    b = nts.FastSampler(graph, seeds, 2, 256, [9, 5], cuda_stream=cs, rng_seed=5)
    a.sample_gpu_fast(256, ssg_id=0, sync=False)
    a.sample_gpu_fast(256, ssg_id=1, sync=False)
    for slot in (0, 1):
        sa = a.wait(slot)
        sb = b.sample_gpu_fast(256)
        for la, lb in zip(sa.sampled_sgs, sb.sampled_sgs):
            assert (la.v_size, la.e_size, la.src_size) == (lb.v_size, lb.e_size, lb.src_size)
            assert torch.equal(la.dev_sample_ans, lb.dev_sample_ans) and torch.equal(la.dev_row_indices, lb.dev_row_indices)
            assert torch.equal(la.dev_source, lb.dev_source) and torch.equal(la.dev_edge_weight_forward, lb.dev_edge_weight_forward)
        assert sa.sampled_sgs[1].dev_row_offset is None and sa.sampled_sgs[0].dev_row_offset is not None
        assert torch.equal(sa.sampled_sgs[0].dev_column_indices, sb.sampled_sgs[0].dev_column_indices)


@pytest.mark.parametrize("name", ["hotness_synth600_l2", "hotness_synth300_l3"])
def test_hotness_pre_sampling_matches_reference_record(nts, cs, name, tmp_path):
    """a11 on the GPU vs what the reference's own preSample wrote (tests/golden, 1 thread): counts, ids, and the .bin bytes."""
    import os
    from golden_util import GOLD
    z = np.load(os.path.join(GOLD, name + ".npz"))
    V, batch, pipeline, layers = (int(x) for x in z["meta"])
    graph = nts.FullyRepGraph(cs, V, edge_pairs=z["pairs"])
    counts, ids = nts.preSample(z["seeds"], batch, pipeline, layers, graph, cache_rate=0.8, cuda_stream=cs)
    assert np.array_equal(counts, z["counts"]) and np.array_equal(ids, z["ids"])
    path = str(tmp_path / "x.bin")
    nts.write_pre_sample_file(path, counts, ids)
    assert np.array_equal(np.fromfile(path, dtype=np.uint32), z["bin_file"])
    take, sub = nts.read_pre_sample_file(path, counts.size, of_rate=0.25)
    assert np.array_equal(take, (counts * 0.25).astype(np.uint32)) and sub.size == take.sum()


def test_hotness_large_graph_and_cache_index(nts, cs):
    V = 30000
    pairs, graph = make_graph(nts, cs, V, 40, seed=31)
    co, ri = oracle.build_csc(pairs, V)
    seeds = np.random.default_rng(8).permutation(V)[:4096].astype(np.uint32)
    for layers, rate in ((2, 0.1), (3, 0.01), (1, 0.5)):
        counts, ids = nts.preSample(seeds, 1024, 2, layers, graph, cache_rate=rate, cuda_stream=cs)
        rc, rids = oracle.pre_sample(seeds, 1024, 2, co, ri, V, layers, cache_rate=rate)
        assert np.array_equal(counts, rc) and np.array_equal(ids, rids), (layers, rate)
    cmap = torch.full((V,), -1, dtype=torch.int32, device="cuda")
    cloc = torch.zeros(V, dtype=torch.int32, device="cuda")
    hot = torch.from_numpy(ids[:counts[0]].view(np.int32)).cuda()
    nts.set_cache_index(cs, cmap, cloc, 7, hot, hot.numel())
    m, l = np.full(V, 0xFFFFFFFF, np.uint32), np.zeros(V, np.uint32)
    oracle.set_cache_index(m, l, 7, ids[:counts[0]])
    assert np.array_equal(u32(cmap), m) and np.array_equal(u32(cloc), l)


# ---------------------------------------------------------------------------------------------
def _power_law_graph_on_gpu(V, E, seed):
    """in-edge CSC generated on the device: power-law in-degree, skewed sources (same recipe as bench.py, scaled)"""
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
    return co.to(torch.int32).cpu().numpy().view(np.uint32), src.cpu().numpy().view(np.uint32)


def test_full_size_products_shaped_cached_path(nts, cs):
    """BASELINE.json configs[2] shape (2.45M vertices, ~62M edges, F=100): hotness -> cache index -> omit sampling -> cached gather
    -> hot-row override, checked through invariants and against torch formulations of the same selections."""
    V, E, F = 2449029, 61859140, 100
    co, ri = _power_law_graph_on_gpu(V, E, 0xBEEF)
    graph = nts.FullyRepGraph(cs, V, column_offset=co, row_indices=ri)
    rng = np.random.default_rng(12)
    train = rng.permutation(V)[:8192].astype(np.uint32)
    counts, ids = nts.preSample(train, 1024, 4, 2, graph, cache_rate=0.05, cuda_stream=cs)   # 2 super-batches
    assert counts.size == 2 and counts.sum() == ids.size and (np.diff(ids[:counts[0]].astype(np.int64)) > 0).all()
    # hottest vertices really are the most frequent in-neighbours of the super-batch (fp64 torch formulation)
    deg = np.diff(co.astype(np.int64))
    sb = train[:4096]
    nbr = np.concatenate([ri[co[d]:co[d + 1]] for d in sb])
    cnt = np.bincount(nbr, minlength=V)
    nnz = int((cnt > 0).sum())
    assert counts[0] == int(np.float32(nnz + 1) * np.float32(0.05))
    pivot = np.sort(cnt)[::-1][counts[0]]
    assert np.array_equal(ids[:counts[0]], np.nonzero(cnt >= pivot)[0][:counts[0]].astype(np.uint32))
    cmap = torch.full((V,), -1, dtype=torch.int32, device="cuda")
    cloc = torch.zeros(V, dtype=torch.int32, device="cuda")
    hot = torch.from_numpy(ids[:counts[0]].view(np.int32)).cuda()
    nts.set_cache_index(cs, cmap, cloc, 0, hot, hot.numel())
    sampler = nts.FastSampler(graph, sb, 2, 1024, [25, 10], cuda_stream=cs)
    sg = sampler.sample_gpu_fast_omit(1024, cmap, 0)
    top, bot = sg.sampled_sgs
    lens = np.diff(u32(bot.dev_column_offset).astype(np.int64))
    dstb = u32(bot.dev_destination)
    is_hot = np.isin(dstb, ids[:counts[0]])
    assert (lens[is_hot] == 0).all() and np.array_equal(lens[~is_hot], np.minimum(deg[dstb[~is_hot]], 10))
    assert is_hot.sum() > 0
    # cached gather: hot rows from the HBM cache table (slot = cache_location), cold rows from the full table
    table = torch.randn((V, F), device="cuda")
    cache_table = table[hot.long()] + 100.0
    hashmap = torch.full((V,), -1, dtype=torch.int32, device="cuda")
    hashmap[hot.long()] = torch.arange(hot.numel(), dtype=torch.int32, device="cuda")
    x = torch.empty((bot.src_size, F), device="cuda")
    hits = torch.zeros(1, dtype=torch.int32, device="cuda")
    sampler.load_feature_gpu_cache(cs, sg, x, table, cache_table, hashmap, hits)
    srcl = bot.dev_source.long()
    ref = torch.where((hashmap[srcl] >= 0)[:, None], table[srcl] + 100.0, table[srcl])
    assert torch.equal(x, ref) and int(hits.item()) == int((hashmap[srcl] >= 0).sum())
    # hot-row override of the aggregated output
    y = nts.SingleGPUAllSampleGraphOp(sg, 1, cs).forward(x)
    assert float(y[torch.from_numpy(is_hot).cuda()].abs().sum()) == 0.0        # omitted columns aggregate to zero rows
    share = torch.randn((hot.numel(), F), device="cuda")
    sampler.load_share_embedding(cs, sg, y, share, cmap, cloc, 0)
    hot_rows = torch.from_numpy(np.nonzero(is_hot)[0]).cuda()
    assert torch.equal(y[hot_rows], share[cloc[bot.dev_destination.long()[hot_rows]].long()])


def test_full_size_papers100m_shaped_sampling(nts, cs):
    """BASELINE.json configs[4] topology scale (111M vertices, ~1.6B edges): u32 offsets up to 1.6e9, a 3.5M-word dedup bitmap,
    1,700 scan tiles. Sampling invariants only (the 57 GB feature table is the sharded-table test's business)."""
    free, _ = torch.cuda.mem_get_info()
    if free < 40 * 2 ** 30:
        pytest.skip("needs ~40 GB of free HBM")
    V, E = 111059956, 1615685872
    co, ri = _power_law_graph_on_gpu(V, E, 0xFACE)
    torch.cuda.empty_cache()
    assert int(co[-1]) == ri.size and ri.size > 1_500_000_000
    graph = nts.FullyRepGraph(cs, V, column_offset=co, row_indices=ri)
    rng = np.random.default_rng(4)
    seeds = rng.integers(V // 2, V, 1024).astype(np.uint32)      # high ids: column offsets beyond 2^30
    seeds = np.unique(seeds)
    sampler = nts.FastSampler(graph, seeds, 2, seeds.size, [25, 10], cuda_stream=cs)
    sg = sampler.sample_gpu_fast(seeds.size)
    deg = np.diff(co.astype(np.int64))
    for lay, f in zip(sg.sampled_sgs, (25, 10)):
        dst = u32(lay.dev_destination)
        lens = np.diff(u32(lay.dev_column_offset).astype(np.int64))
        assert np.array_equal(lens, np.minimum(deg[dst], f))
        src = u32(lay.dev_source).astype(np.int64)
        ans = u32(lay.dev_sample_ans)
        assert (np.diff(src) > 0).all() and np.array_equal(np.unique(ans), src)
        assert np.array_equal(src[u32(lay.dev_row_indices)], ans)
        lco = u32(lay.dev_column_offset)
        for j in range(0, dst.size, max(1, dst.size // 300)):       # membership on a sample of columns
            nb = ri[co[dst[j]]:co[dst[j] + 1]]
            got = ans[lco[j]:lco[j + 1]]
            assert np.isin(got, nb).all()
        ro = u32(lay.dev_row_offset)
        assert ro[-1] == lay.e_size
    assert torch.equal(sg.sampled_sgs[1].dev_destination, sg.sampled_sgs[0].dev_source)


def test_cold_row_staging_overlapped_slots(nts, cs):
    """Hot rows from the HBM cache table, cold rows staged from (pageable) host memory by the worker thread through pinned
    memory + one async copy on a side stream; two slots in flight; result identical to the oracle's cached gather."""
    V, F = 40000, 100
    rng = np.random.default_rng(6)
    table_np = rng.standard_normal((V, F)).astype(np.float32)
    hot = np.sort(rng.permutation(V)[:6000]).astype(np.uint32)
    cache_np = table_np[hot] + np.float32(50.0)
    hashmap = np.full(V, 0xFFFFFFFF, np.uint32)
    hashmap[hot] = np.arange(hot.size, dtype=np.uint32)
    d_hash = torch.from_numpy(hashmap.view(np.int32)).cuda()
    d_cache = torch.from_numpy(cache_np).cuda()
    host_table = torch.from_numpy(table_np)                          # plain pageable host memory
    stage = nts.ColdStage(cs, host_table, max_rows=30000)
    batches = [rng.integers(0, V, n).astype(np.uint32) for n in (30000, 1, 17000, 0, 25000)]
    d_ids = [torch.from_numpy(b.view(np.int32)).cuda() for b in batches]
    outs = [torch.full((max(b.size, 1), F), 9.0, device="cuda") for b in batches]
    stage.submit(0, d_ids[0], batches[0].size, d_hash)
    for i, b in enumerate(batches):
        if i + 1 < len(batches):
            stage.submit((i + 1) % 2, d_ids[i + 1], batches[i + 1].size, d_hash)   # next batch stages while this one is merged
        n_cold = stage.gather(i % 2, outs[i], d_cache, d_hash, d_ids[i])
        assert n_cold == int((hashmap[b] == 0xFFFFFFFF).sum())
    cs.CUDA_DEVICE_SYNCHRONIZE()
    for b, o in zip(batches, outs):
        if b.size:
            assert np.array_equal(bits(f32(o)[:b.size]), bits(oracle.gather_rows_cached(table_np, cache_np, hashmap, b)))
    with pytest.raises(nts.NtsError):
        stage.submit(0, d_ids[0], 30001, d_hash)
    # the same with the hot cache partitioned over three shards (cache slot k -> shard k % 3, row k // 3), pitched rows
    pitch = F + 4
    shards = []
    for r in range(3):
        t = torch.zeros((len(range(r, hot.size, 3)), pitch), device="cuda")
        t[:, :F] = torch.from_numpy(cache_np[r::3]).cuda()
        shards.append(t)
    hot_table = nts.FeatureTable(cs, shards, F, pitch, hot.size, keepalive=shards)
    outs2 = [torch.full((max(b.size, 1), F), 7.0, device="cuda") for b in batches]
    stage.submit(0, d_ids[0], batches[0].size, d_hash)
    for i, b in enumerate(batches):
        if i + 1 < len(batches):
            stage.submit((i + 1) % 2, d_ids[i + 1], batches[i + 1].size, d_hash)
        assert stage.gather_table(i % 2, outs2[i], hot_table, d_hash, d_ids[i]) == int((hashmap[b] == 0xFFFFFFFF).sum())
    cs.CUDA_DEVICE_SYNCHRONIZE()
    for b, o in zip(batches, outs2):
        if b.size:
            assert np.array_equal(bits(f32(o)[:b.size]), bits(oracle.gather_rows_cached(table_np, cache_np, hashmap, b)))
