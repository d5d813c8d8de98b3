"""ctypes/numpy front end of oracle/oracle.c (TEST INFRASTRUCTURE ONLY).

Each function forwards to the C restatement, whose comments cite the reference file:line it
follows. `build()` compiles liboracle.so with the flags that define the oracle's arithmetic
(-ffp-contract=off: multiply then add, as core/ntsBaseOp.hpp:546-562 spells it).
"""
import ctypes
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB = None
U32 = np.uint32
F32 = np.float32


def build(force=False):
    so = os.path.join(_DIR, "liboracle.so")
    src = os.path.join(_DIR, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["/usr/bin/gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off",
                               "-o", so, src, "-lm"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _u32(a):
    return np.ascontiguousarray(a, dtype=U32)


def _f32(a):
    return np.ascontiguousarray(a, dtype=F32)


def _call(name, restype, *args):
    fn = getattr(lib(), name)
    fn.restype = restype
    conv = []
    for a in args:
        if isinstance(a, np.ndarray) or a is None:
            conv.append(_p(a))
        else:
            conv.append(a)
    return fn(*conv)


c_u32 = ctypes.c_uint32
c_u64 = ctypes.c_uint64
c_int = ctypes.c_int


def build_csc(pairs, n_vertices):
    pairs = _u32(pairs).reshape(-1, 2)
    E = pairs.shape[0]
    co = np.zeros(n_vertices + 1, U32)
    ri = np.zeros(max(E, 1), U32)
    _call("orc_build_csc", None, pairs, c_u64(E), c_u32(n_vertices), co, ri)
    return co, ri[:E]


def degrees(pairs, n_vertices):
    pairs = _u32(pairs).reshape(-1, 2)
    ind = np.zeros(n_vertices, U32)
    outd = np.zeros(n_vertices, U32)
    _call("orc_degrees", None, pairs, c_u64(pairs.shape[0]), c_u32(n_vertices), ind, outd)
    return ind, outd


def count_offsets(destination, g_column_offset, fanout, skip=None, skip_value=0xFFFFFFFF):
    destination = _u32(destination)
    co = np.zeros(destination.size + 1, U32)
    sk = None if skip is None else _u32(skip)
    e = _call("orc_count_offsets", c_u32, destination, c_u32(destination.size), _u32(g_column_offset), c_int(fanout),
              sk, c_u32(skip_value), co)
    return co, int(e)


def sample_layer(destination, column_offset, g_column_offset, g_row_indices, fanout, seed):
    destination = _u32(destination)
    column_offset = _u32(column_offset)
    E = int(column_offset[-1])
    ans = np.zeros(max(E, 1), U32)
    _call("orc_sample_layer", None, destination, c_u32(destination.size), column_offset, _u32(g_column_offset),
          _u32(g_row_indices), c_int(fanout), c_u64(seed), ans)
    return ans[:E]


def reindex(sample_ans, destination, n_vertices, merge_src_dst=False):
    sample_ans = _u32(sample_ans)
    destination = _u32(destination)
    E = sample_ans.size
    source = np.zeros(E + destination.size + 1, U32)
    ri = np.zeros(max(E, 1), U32)
    dl = np.zeros(max(destination.size, 1), U32) if merge_src_dst else None
    s = _call("orc_reindex", c_u32, sample_ans, c_u32(E), destination, c_u32(destination.size), c_u32(n_vertices),
              c_int(1 if merge_src_dst else 0), source, ri, dl)
    return source[:s].copy(), ri[:E], (dl[:destination.size] if merge_src_dst else None)


def csc_to_csr(column_offset, row_indices, n_src):
    column_offset = _u32(column_offset)
    row_indices = _u32(row_indices)
    V = column_offset.size - 1
    E = row_indices.size
    ro = np.zeros(n_src + 1, U32)
    ci = np.zeros(max(E, 1), U32)
    _call("orc_csc_to_csr", None, column_offset, row_indices, c_u32(V), c_u32(n_src), c_u32(E), ro, ci)
    return ro, ci[:E]


def update_degrees(destination, source, column_offset, row_indices, n_vertices):
    ind = np.zeros(n_vertices, U32)
    outd = np.zeros(n_vertices, U32)
    destination = _u32(destination)
    _call("orc_update_degrees", None, destination, _u32(source), _u32(column_offset), _u32(row_indices),
          c_u32(destination.size), c_u32(n_vertices), ind, outd)
    return ind, outd


def weights(destination, source, column_offset, row_indices, row_offset, column_indices, in_degree, out_degree,
            weight_type=0):
    destination = _u32(destination)
    source = _u32(source)
    E = int(_u32(column_offset)[-1])
    ewf = np.zeros(max(E, 1), F32)
    ewb = np.zeros(max(E, 1), F32) if row_offset is not None else None
    _call("orc_weights", None, destination, source, _u32(column_offset), _u32(row_indices),
          None if row_offset is None else _u32(row_offset), None if column_indices is None else _u32(column_indices),
          c_u32(destination.size), c_u32(source.size), _u32(in_degree), _u32(out_degree), c_int(weight_type), ewf, ewb)
    return ewf[:E], (None if ewb is None else ewb[:E])


def gather_rows(table, ids):
    table = _f32(table)
    ids = _u32(ids)
    F = table.shape[1]
    out = np.zeros((ids.size, F), F32)
    _call("orc_gather_rows", None, table, ids, c_u32(ids.size), c_u32(F), out)
    return out


def gather_labels(labels, ids):
    labels = np.ascontiguousarray(labels, dtype=np.int64)
    ids = _u32(ids)
    out = np.zeros(ids.size, np.int64)
    _call("orc_gather_labels", None, labels, ids, c_u32(ids.size), out)
    return out


def gather_rows_cached(full_table, cache_table, cache_node_hashmap, ids):
    full_table = _f32(full_table)
    ids = _u32(ids)
    F = full_table.shape[1]
    out = np.zeros((ids.size, F), F32)
    _call("orc_gather_rows_cached", None, full_table, _f32(cache_table), _u32(cache_node_hashmap), ids,
          c_u32(ids.size), c_u32(F), out)
    return out


def row_override(out, share, cache_map, cache_location, destination, super_batch_id):
    out = _f32(out).copy()
    destination = _u32(destination)
    _call("orc_row_override", None, out, _f32(share), _u32(cache_map), _u32(cache_location), destination,
          c_u32(destination.size), c_u32(out.shape[1]), c_u32(super_batch_id))
    return out


def aggregate_fwd(x, column_offset, row_indices, weight=None, destination=None, source=None, in_degree=None,
                  out_degree=None):
    x = _f32(x)
    column_offset = _u32(column_offset)
    V = column_offset.size - 1
    F = x.shape[1]
    y = np.zeros((V, F), F32)
    _call("orc_aggregate_fwd", None, x, y, None if weight is None else _f32(weight), column_offset, _u32(row_indices),
          None if destination is None else _u32(destination), None if source is None else _u32(source),
          None if in_degree is None else _u32(in_degree), None if out_degree is None else _u32(out_degree),
          c_u32(V), c_u32(F))
    return y


def aggregate_bwd(dy, column_offset, row_indices, n_src, weight=None, destination=None, source=None, in_degree=None,
                  out_degree=None):
    dy = _f32(dy)
    column_offset = _u32(column_offset)
    V = column_offset.size - 1
    F = dy.shape[1]
    dx = np.zeros((n_src, F), F32)
    _call("orc_aggregate_bwd", None, dy, dx, None if weight is None else _f32(weight), column_offset, _u32(row_indices),
          None if destination is None else _u32(destination), None if source is None else _u32(source),
          None if in_degree is None else _u32(in_degree), None if out_degree is None else _u32(out_degree),
          c_u32(V), c_u32(n_src), c_u32(F))
    return dx


def aggregate_bwd_csr(dy, row_offset, column_indices, weight_b=None):
    dy = _f32(dy)
    row_offset = _u32(row_offset)
    S = row_offset.size - 1
    F = dy.shape[1]
    dx = np.zeros((S, F), F32)
    _call("orc_aggregate_bwd_csr", None, dy, dx, None if weight_b is None else _f32(weight_b), row_offset,
          _u32(column_indices), c_u32(S), c_u32(F))
    return dx


def scatter_src_dst(x, column_offset, row_indices, dst_local_id):
    x = _f32(x)
    column_offset = _u32(column_offset)
    V = column_offset.size - 1
    F = x.shape[1]
    E = int(column_offset[-1])
    msg = np.zeros((E, 2 * F), F32)
    _call("orc_scatter_src_dst", None, x, msg, column_offset, _u32(row_indices), _u32(dst_local_id), c_u32(V), c_u32(F))
    return msg


def gather_src_dst(dmsg, column_offset, row_indices, dst_local_id, n_src):
    dmsg = _f32(dmsg)
    column_offset = _u32(column_offset)
    V = column_offset.size - 1
    F = dmsg.shape[1] // 2
    dx = np.zeros((n_src, F), F32)
    _call("orc_gather_src_dst", None, dx, dmsg, column_offset, _u32(row_indices), _u32(dst_local_id), c_u32(V),
          c_u32(n_src), c_u32(F))
    return dx


def edge_softmax_fwd(m, column_offset):
    m = _f32(m).reshape(-1)
    column_offset = _u32(column_offset)
    a = np.zeros_like(m)
    _call("orc_edge_softmax_fwd", None, m, a, column_offset, c_u32(column_offset.size - 1))
    return a


def edge_softmax_bwd(da, a, column_offset):
    da = _f32(da).reshape(-1)
    a = _f32(a).reshape(-1)
    column_offset = _u32(column_offset)
    dm = np.zeros_like(a)
    _call("orc_edge_softmax_bwd", None, da, a, dm, column_offset, c_u32(column_offset.size - 1))
    return dm


def gather_msg_to_dst(msg, column_offset):
    msg = _f32(msg)
    column_offset = _u32(column_offset)
    V = column_offset.size - 1
    y = np.zeros((V, msg.shape[1]), F32)
    _call("orc_gather_msg_to_dst", None, y, msg, column_offset, c_u32(V), c_u32(msg.shape[1]))
    return y


def scatter_dst_to_msg(y, column_offset):
    y = _f32(y)
    column_offset = _u32(column_offset)
    V = column_offset.size - 1
    msg = np.zeros((int(column_offset[-1]), y.shape[1]), F32)
    _call("orc_scatter_dst_to_msg", None, msg, y, column_offset, c_u32(V), c_u32(y.shape[1]))
    return msg


def gat_layer_fwd(h, att, column_offset, row_indices, dst_local_id):
    h = _f32(h)
    column_offset = _u32(column_offset)
    V = column_offset.size - 1
    F = h.shape[1]
    E = int(column_offset[-1])
    pre = np.zeros(max(E, 1), F32)
    alpha = np.zeros(max(E, 1), F32)
    out = np.zeros((V, F), F32)
    _call("orc_gat_layer_fwd", None, h, _f32(att).reshape(-1), column_offset, _u32(row_indices), _u32(dst_local_id),
          c_u32(V), c_u32(F), pre, alpha, out)
    return out, alpha[:E], pre[:E]


def gat_layer_bwd(h, att, dout, score_pre, alpha, column_offset, row_indices, dst_local_id):
    h = _f32(h)
    column_offset = _u32(column_offset)
    V = column_offset.size - 1
    S, F = h.shape
    dh = np.zeros((S, F), F32)
    datt = np.zeros(2 * F, F32)
    _call("orc_gat_layer_bwd", None, h, _f32(att).reshape(-1), _f32(dout), _f32(score_pre), _f32(alpha), column_offset,
          _u32(row_indices), _u32(dst_local_id), c_u32(V), c_u32(S), c_u32(F), dh, datt)
    return dh, datt


def set_cache_index(cache_map, cache_location, super_batch_id, cache_ids):
    cache_ids = _u32(cache_ids)
    _call("orc_set_cache_index", None, cache_map, cache_location, c_u32(super_batch_id), cache_ids, c_u32(cache_ids.size))


def feat(v, j):
    """The formula-defined synthetic feature table of oracle/ref_driver.cpp (feat())."""
    v = np.asarray(v, dtype=np.uint64)[:, None]
    j = np.asarray(j, dtype=np.uint64)[None, :]
    m = np.uint64(0xFFFFFFFF)
    h = (v * np.uint64(2654435761) + j * np.uint64(40503) + np.uint64(12345)) & m
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(2246822519)) & m
    h ^= h >> np.uint64(13)
    return ((h % np.uint64(257)).astype(np.int64) - 128).astype(np.float32) / np.float32(64.0)


def sample_batch(seeds, g_column_offset, g_row_indices, fanouts, n_vertices, in_degree, out_degree, seed=0,
                 weight_type=0, up_degree=False, merge_src_dst=False, replay=None):
    """All layers of one mini-batch, in the order of FastSampler::sample_fast
    (core/ntsFastSampler.hpp:962-1140). `replay`, if given, is a list of per-layer sample_ans
    arrays to use instead of drawing (the bit-exact replay path). Returns a list of dicts."""
    layers = []
    dst = _u32(seeds)
    ind, outd = _u32(in_degree), _u32(out_degree)
    for i, f in enumerate(fanouts):
        co, E = count_offsets(dst, g_column_offset, f)
        ans = _u32(replay[i]) if replay is not None else sample_layer(dst, co, g_column_offset, g_row_indices, f,
                                                                       seed + 1000003 * i)
        src, ri, dl = reindex(ans, dst, n_vertices, merge_src_dst)
        ro, ci = csc_to_csr(co, ri, src.size)
        if up_degree:
            ind, outd = update_degrees(dst, src, co, ri, n_vertices)
        ewf, ewb = (None, None) if weight_type is None else weights(dst, src, co, ri, ro, ci, ind, outd, weight_type)
        layers.append(dict(destination=dst.copy(), column_offset=co, sample_ans=ans, source=src, row_indices=ri,
                           row_offset=ro, column_indices=ci, e_w_f=ewf, e_w_b=ewb, dst_local_id=dl,
                           in_deg=ind.copy() if up_degree else None, out_deg=outd.copy() if up_degree else None))
        dst = src
    return layers


def hotness(seeds, g_column_offset, g_row_indices, n_vertices, layers, cache_rate):
    """one super-batch of core/ntsBaseOp.hpp:333-399 -> (hot ids ascending, final counts)"""
    seeds = _u32(seeds)
    ids = np.zeros(n_vertices, U32)
    counts = np.zeros(n_vertices, U32)
    n = _call("orc_hotness", c_u32, seeds, c_u32(seeds.size), _u32(g_column_offset), _u32(g_row_indices), c_u32(n_vertices),
              c_int(layers), ctypes.c_float(cache_rate), ids, counts)
    return ids[:n].copy(), counts


def pre_sample(train_ids, batch_size, pipeline_num, g_column_offset, g_row_indices, n_vertices, layers, cache_rate=0.8):
    """nts::op::preSample (core/ntsBaseOp.hpp:415-470): per super-batch (= batch_size * pipeline_num seeds) hot lists."""
    train_ids = _u32(train_ids)
    sb = batch_size * pipeline_num
    counts, ids = [], []
    for start in range(0, train_ids.size, sb):
        h, _ = hotness(train_ids[start:start + sb], g_column_offset, g_row_indices, n_vertices, layers, cache_rate)
        counts.append(h.size)
        ids.append(h)
    return np.array(counts, U32), (np.concatenate(ids) if ids else np.zeros(0, U32))


def pre_sample_file_pack(counts, ids):
    """on-disk layout (core/ntsBaseOp.hpp:477-495): u32 counts[#super_batches] || u32 ids[sum(counts)]"""
    return np.concatenate([_u32(counts), _u32(ids)])


def pre_sample_file_unpack(raw, n_super_batches, of_rate=1.0):
    """reader (core/ntsBaseOp.hpp:497-538): the first counts[i]*of_rate ids of every group"""
    raw = _u32(raw)
    counts = raw[:n_super_batches]
    take = (counts.astype(np.float32) * np.float32(of_rate)).astype(U32)
    out, pos = [], n_super_batches
    for c, t in zip(counts, take):
        out.append(raw[pos:pos + t])
        pos += int(c)
    return take, (np.concatenate(out) if out else np.zeros(0, U32))
