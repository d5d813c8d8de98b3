/* oracle/oracle.c -- CPU restatement of the reference's sample-based hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is imported, linked or executed by the
 * product (sample-based-gnn_b200/, include/): only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use it, and only as the checker.
 *
 * Parity status: PINNED. Every function below is checked (tests/test_oracle_golden.py)
 * against outputs of the reference's own code (oracle/_ref/ref_driver, built from the sources
 * under /root/reference by oracle/Makefile) recorded into tests/golden/ by
 * oracle/make_golden.py. Integer outputs match bit for bit; fp32 aggregation matches bit for
 * bit when the reference is compiled with -ffp-contract=off and OMP_NUM_THREADS=1.
 *
 * Paths are relative to /root/reference. Types follow dep/gemini/type.hpp:29-31
 * (VertexId = uint32_t, ValueType = float).
 *
 * Plain C11, single-threaded, no dependencies: `gcc -O2 -ffp-contract=off -shared -fPIC`.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint32_t vid_t;

/* ------------------------------------------------------------------------------------------
 * Global graph: in-edge CSC from an edge list of (src,dst) u32 pairs.
 * core/FullyRepGraph.hpp:724-798 (ReadRepGraphFromRawFile): pass 1 counts tmp_offset[dst+1]++,
 * inclusive scan, pass 2 row_indices[tmp_offset[dst]++] = src => edge order inside a column
 * is file order. */
void orc_build_csc(const vid_t *pairs, uint64_t n_edges, vid_t n_vertices, vid_t *column_offset /*[V+1]*/,
                   vid_t *row_indices /*[E]*/) {
  vid_t *tmp = (vid_t *)calloc((size_t)n_vertices + 1, sizeof(vid_t));
  for (uint64_t e = 0; e < n_edges; e++) tmp[pairs[2 * e + 1] + 1]++;
  for (vid_t i = 0; i < n_vertices; i++) tmp[i + 1] += tmp[i];
  memcpy(column_offset, tmp, sizeof(vid_t) * ((size_t)n_vertices + 1));
  for (uint64_t e = 0; e < n_edges; e++) row_indices[tmp[pairs[2 * e + 1]]++] = pairs[2 * e];
  free(tmp);
}

/* Degrees used by the edge weights: per-edge counts (core/graph.hpp:1305-1424), clamped to
 * >= 1 (core/graph.hpp:4525-4530). */
void orc_degrees(const vid_t *pairs, uint64_t n_edges, vid_t n_vertices, vid_t *in_degree, vid_t *out_degree) {
  memset(in_degree, 0, sizeof(vid_t) * n_vertices);
  memset(out_degree, 0, sizeof(vid_t) * n_vertices);
  for (uint64_t e = 0; e < n_edges; e++) { out_degree[pairs[2 * e]]++; in_degree[pairs[2 * e + 1]]++; }
  for (vid_t v = 0; v < n_vertices; v++) { if (in_degree[v] < 1) in_degree[v] = 1; if (out_degree[v] < 1) out_degree[v] = 1; }
}

/* ------------------------------------------------------------------------------------------
 * Per-layer column offsets: serial exclusive scan of min(deg, fanout).
 * core/FullyRepGraph.hpp:530-539 (init_co_only) with the count lambda of
 * core/ntsFastSampler.hpp:1001-1009: ret = min((int)nbrs, fanout); ret == -1 -> nbrs (take all).
 * `skip` (may be NULL) restates the GPU omit variants: a dst whose flag equals `skip_value`
 * contributes 0 edges (cuda/ntsCUDATransferKernel.cuh:797-822, the super-batch overload); with
 * skip_value == 0xffffffff the test is flag != -1 -> 0 edges (ibid. :771-795).
 * Returns the edge count. */
vid_t orc_count_offsets(const vid_t *destination, vid_t n_dst, const vid_t *g_column_offset, int fanout,
                        const vid_t *skip, vid_t skip_value, vid_t *column_offset /*[n_dst+1]*/) {
  vid_t off = 0;
  for (vid_t i = 0; i < n_dst; i++) {
    column_offset[i] = off;
    vid_t d = destination[i];
    vid_t nbrs = g_column_offset[d + 1] - g_column_offset[d];
    int r = (int)nbrs < fanout ? (int)nbrs : fanout;
    vid_t ret = (vid_t)r;
    if (r == -1) ret = nbrs;
    if (skip) {
      if (skip_value == 0xffffffffu) { if (skip[d] != 0xffffffffu) ret = 0; }
      else if (skip[d] == skip_value) ret = 0;
    }
    off += ret;
  }
  column_offset[n_dst] = off;
  return off;
}

/* splitmix64: the oracle sampler's own generator (the reference's is a thread_local
 * mt19937(2000) whose draw-to-dst assignment depends on OpenMP scheduling,
 * core/ntsFastSampler.hpp:200-205 -> only the distribution is a contract). */
static uint64_t splitmix64(uint64_t *s) {
  uint64_t z = (*s += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

/* Neighbour selection for one layer. core/ntsFastSampler.hpp:1020-1054:
 *   nbr_size > fanout : draw uniform offsets in [0,nbr_size) until `num` distinct ones are held
 *                       (rejection into a set) -> uniform `num`-subset, no duplicates;
 *   otherwise         : all in-neighbours in stored order.
 * Emits global source ids into sample_ans[column_offset[i] ...]. The order inside a sampled
 * column is insertion order here (the reference's is libstdc++ unordered_map iteration order,
 * which is not a contract). */
void orc_sample_layer(const vid_t *destination, vid_t n_dst, const vid_t *column_offset, const vid_t *g_column_offset,
                      const vid_t *g_row_indices, int fanout, uint64_t seed, vid_t *sample_ans) {
  for (vid_t i = 0; i < n_dst; i++) {
    vid_t d = destination[i];
    vid_t base = g_column_offset[d];
    vid_t nbr = g_column_offset[d + 1] - base;
    vid_t num = column_offset[i + 1] - column_offset[i];
    vid_t *out = sample_ans + column_offset[i];
    if (fanout >= 0 && nbr > (vid_t)fanout && num > 0) {
      /* independent stream per dst slot: hash (seed, i) first -- consecutive splitmix64 states differ by
       * the golden-ratio increment, so seeding with seed + i*increment would make the streams shifted copies */
      uint64_t t = seed ^ ((uint64_t)i * 0xd6e8feb86659fd93ull);
      uint64_t s = splitmix64(&t);
      s ^= splitmix64(&t) << 1;
      vid_t have = 0;
      vid_t *pos = (vid_t *)malloc(sizeof(vid_t) * num);
      while (have < num) {
        vid_t r = (vid_t)(splitmix64(&s) % nbr);
        int dup = 0;
        for (vid_t k = 0; k < have; k++) if (pos[k] == r) { dup = 1; break; }
        if (!dup) pos[have++] = r;
      }
      for (vid_t k = 0; k < num; k++) out[k] = g_row_indices[base + pos[k]];
      free(pos);
    } else {
      for (vid_t k = 0; k < num; k++) out[k] = g_row_indices[base + k];
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * Dedup + reindex. core/ntsFastSampler.hpp:1017,1037,1046,1050-1052 (bitmap marks, plus every
 * dst when is_merge_src_dst), :1062-1083 (serial scan of the bitmap words, ascending id ->
 * source[], src_index_array[v] = rank), :1085-1093 (row_indices[e] = src_index_array[sample_ans[e]]).
 * Bitmap layout: dep/gemini/bitmap.hpp:7-8,58-64 (64-bit words, bit i&63).
 * dst_local_id (GAT): the reference writes src_index_array[dst()[layer]] (its line :1096 indexes
 * by the layer number -- SURVEY.md section 8 lists this as a quirk NOT to copy); the intended
 * value, which the GPU path computes (cuda/ntsCUDATransferKernel.cuh:1189-1196), is the local
 * id of destination[i] and that is what is restated here.
 * Returns src_size. */
vid_t orc_reindex(const vid_t *sample_ans, vid_t n_edges, const vid_t *destination, vid_t n_dst, vid_t n_vertices,
                  int merge_src_dst, vid_t *source /*[<= n_edges (+n_dst)]*/, vid_t *row_indices /*[n_edges]*/,
                  vid_t *dst_local_id /*[n_dst] or NULL*/) {
  size_t words = ((size_t)n_vertices >> 6) + 1;
  uint64_t *bm = (uint64_t *)calloc(words, sizeof(uint64_t));
  vid_t *src_index = (vid_t *)malloc(sizeof(vid_t) * ((size_t)n_vertices + 64));
  for (vid_t e = 0; e < n_edges; e++) bm[sample_ans[e] >> 6] |= 1ull << (sample_ans[e] & 63);
  if (merge_src_dst) for (vid_t i = 0; i < n_dst; i++) bm[destination[i] >> 6] |= 1ull << (destination[i] & 63);
  vid_t n_src = 0;
  for (size_t base = 0; base < n_vertices; base += 64) {
    uint64_t w = bm[base >> 6];
    vid_t off = 0;
    while (w) {
      if (w & 1) { src_index[base + off] = n_src; source[n_src] = (vid_t)(base + off); n_src++; }
      off++;
      w >>= 1;
    }
  }
  for (vid_t e = 0; e < n_edges; e++) row_indices[e] = src_index[sample_ans[e]];
  if (merge_src_dst && dst_local_id) for (vid_t i = 0; i < n_dst; i++) dst_local_id[i] = src_index[destination[i]];
  free(bm);
  free(src_index);
  return n_src;
}

/* CSC -> CSR. core/coocsc.hpp:82-111: histogram of row_indices, exclusive scan, stable fill in
 * (dst ascending, CSC position ascending) order, shift back. The reference's histogram loop is
 * a racy `omp parallel for` (:87-90); the single-threaded result is the contract. */
void orc_csc_to_csr(const vid_t *column_offset, const vid_t *row_indices, vid_t n_dst, vid_t n_src, vid_t n_edges,
                    vid_t *row_offset /*[n_src+1]*/, vid_t *column_indices /*[n_edges]*/) {
  memset(row_offset, 0, sizeof(vid_t) * ((size_t)n_src + 1));
  for (vid_t e = 0; e < n_edges; e++) row_offset[row_indices[e]]++;
  vid_t cum = 0;
  for (vid_t i = 0; i < n_src; i++) { vid_t t = row_offset[i]; row_offset[i] = cum; cum += t; }
  row_offset[n_src] = n_edges;
  for (vid_t i = 0; i < n_dst; i++)
    for (vid_t j = column_offset[i]; j < column_offset[i + 1]; j++) {
      vid_t col = row_indices[j];
      column_indices[row_offset[col]++] = i;
    }
  vid_t last = 0;
  for (vid_t c = 0; c <= n_src; c++) { vid_t t = row_offset[c]; row_offset[c] = last; last = t; }
}

/* UP_DEGREE: per-batch sampled degrees. core/FullyRepGraph.hpp:189-207 (update_degrees): zero
 * both |V| arrays, ins[dst] += column length, outs[src]++ per sampled edge. No clamp. */
void orc_update_degrees(const vid_t *destination, const vid_t *source, const vid_t *column_offset,
                        const vid_t *row_indices, vid_t n_dst, vid_t n_vertices, vid_t *in_degree, vid_t *out_degree) {
  memset(in_degree, 0, sizeof(vid_t) * n_vertices);
  memset(out_degree, 0, sizeof(vid_t) * n_vertices);
  for (vid_t i = 0; i < n_dst; i++) {
    in_degree[destination[i]] += column_offset[i + 1] - column_offset[i];
    for (vid_t j = column_offset[i]; j < column_offset[i + 1]; j++) out_degree[source[row_indices[j]]]++;
  }
}

/* nts_norm_degree, core/ntsBaseOp.hpp:652-657:
 *   1 / ((float)std::sqrt(out_degree[src]) * (float)std::sqrt(in_degree[dst]))
 * std::sqrt(uint32) is the double sqrt; each factor is rounded to float, the product and the
 * reciprocal are fp32. */
static float norm_degree(const vid_t *in_degree, const vid_t *out_degree, vid_t src, vid_t dst) {
  float a = (float)sqrt((double)out_degree[src]);
  float b = (float)sqrt((double)in_degree[dst]);
  float p = a * b;
  return 1.0f / p;
}

/* Edge weights. core/coocsc.hpp:301-324 (WeightCompute: e_w_b in CSR order, e_w_f in CSC order)
 * with the lambdas of core/ntsFastSampler.hpp:1111-1119:
 *   weight_type 0 (Sum)  : nts_norm_degree(src,dst)
 *   weight_type 1 (Mean) : nts_norm_degree(src,dst) / in_degree[dst]        (uint -> float divide)
 *   weight_type 2        : the GPU get_mean_weight variant, / (sampled column length)
 *                          (cuda/ntsCUDATransferKernel.cuh:319-342)
 * Either output may be NULL. */
void orc_weights(const vid_t *destination, const vid_t *source, const vid_t *column_offset, const vid_t *row_indices,
                 const vid_t *row_offset, const vid_t *column_indices, vid_t n_dst, vid_t n_src,
                 const vid_t *in_degree, const vid_t *out_degree, int weight_type, float *e_w_f, float *e_w_b) {
  if (e_w_b)
    for (vid_t i = 0; i < n_src; i++)
      for (vid_t j = row_offset[i]; j < row_offset[i + 1]; j++) {
        vid_t d = column_indices[j];
        float w = norm_degree(in_degree, out_degree, source[i], destination[d]);
        if (weight_type == 1) w = w / (float)in_degree[destination[d]];
        if (weight_type == 2) w = w / (float)(uint64_t)(column_offset[d + 1] - column_offset[d]);
        e_w_b[j] = w;
      }
  if (e_w_f)
    for (vid_t i = 0; i < n_dst; i++)
      for (vid_t j = column_offset[i]; j < column_offset[i + 1]; j++) {
        float w = norm_degree(in_degree, out_degree, source[row_indices[j]], destination[i]);
        if (weight_type == 1) w = w / (float)in_degree[destination[i]];
        if (weight_type == 2) w = w / (float)(uint64_t)(column_offset[i + 1] - column_offset[i]);
        e_w_f[j] = w;
      }
}

/* ------------------------------------------------------------------------------------------
 * Feature / label gather. core/ntsMiniBatchGraphOp.hpp:36-60. */
void orc_gather_rows(const float *table, const vid_t *ids, vid_t n, vid_t feature_size, float *out) {
  for (vid_t i = 0; i < n; i++)
    memcpy(out + (size_t)i * feature_size, table + (size_t)ids[i] * feature_size, sizeof(float) * feature_size);
}
void orc_gather_labels(const int64_t *labels, const vid_t *ids, vid_t n, int64_t *out) {
  for (vid_t i = 0; i < n; i++) out[i] = labels[ids[i]];
}

/* Cached gather. core/ntsFastSampler.hpp:263-317 + cuda/ntsCUDATransferKernel.cuh:154-183:
 * a row whose cache_node_hashmap[id] != -1 comes from the device cache table at that slot,
 * every other row from the full (host) table. */
void orc_gather_rows_cached(const float *full_table, const float *cache_table, const vid_t *cache_node_hashmap,
                            const vid_t *ids, vid_t n, vid_t feature_size, float *out) {
  for (vid_t i = 0; i < n; i++) {
    vid_t slot = cache_node_hashmap[ids[i]];
    const float *src = slot != 0xffffffffu ? cache_table + (size_t)slot * feature_size
                                            : full_table + (size_t)ids[i] * feature_size;
    memcpy(out + (size_t)i * feature_size, src, sizeof(float) * feature_size);
  }
}

/* Hot-row override. cuda/ntsCUDATransferKernel.cuh:412-497: for every dst row i whose
 * cache_map[destination[i]] == super_batch_id, out[i,:] = share[cache_location[destination[i]],:]. */
void orc_row_override(float *out, const float *share, const vid_t *cache_map, const vid_t *cache_location,
                      const vid_t *destination, vid_t n_dst, vid_t feature_size, vid_t super_batch_id) {
  for (vid_t i = 0; i < n_dst; i++) {
    vid_t v = destination[i];
    if (cache_map[v] == super_batch_id)
      memcpy(out + (size_t)i * feature_size, share + (size_t)cache_location[v] * feature_size, sizeof(float) * feature_size);
  }
}

/* ------------------------------------------------------------------------------------------
 * Aggregation forward. core/ntsMiniBatchGraphOp.hpp:153-182: per dst, sequentially over the
 * column in CSC order, nts_comp(out, in, w, F) = out[k] += in[k]*w (mul then add, not fused:
 * core/ntsBaseOp.hpp:546-562). weight == NULL recomputes nts_norm_degree per edge as the CPU op
 * does; otherwise the stored per-edge weights are used (GPU ops, core/ntsSingleGPUSampleGraphOp.hpp:73). */
void orc_aggregate_fwd(const float *x /*[n_src,F]*/, float *y /*[n_dst,F]*/, const float *weight,
                       const vid_t *column_offset, const vid_t *row_indices, const vid_t *destination, const vid_t *source,
                       const vid_t *in_degree, const vid_t *out_degree, vid_t n_dst, vid_t feature_size) {
  memset(y, 0, sizeof(float) * (size_t)n_dst * feature_size);
  for (vid_t d = 0; d < n_dst; d++) {
    float *o = y + (size_t)d * feature_size;
    for (vid_t e = column_offset[d]; e < column_offset[d + 1]; e++) {
      vid_t ls = row_indices[e];
      float w = weight ? weight[e] : norm_degree(in_degree, out_degree, source[ls], destination[d]);
      const float *in = x + (size_t)ls * feature_size;
      for (vid_t k = 0; k < feature_size; k++) { float t = in[k] * w; o[k] = o[k] + t; }
    }
  }
}

/* Aggregation backward. core/ntsMiniBatchGraphOp.hpp:214-268: dX[r_i[e],:] += w * dY[d,:]
 * iterating dst ascending then CSC order (nts_acc, core/ntsBaseOp.hpp:579-584; the reference
 * uses CAS atomics across OpenMP threads -- the single-threaded order is restated). */
void orc_aggregate_bwd(const float *dy /*[n_dst,F]*/, float *dx /*[n_src,F]*/, const float *weight,
                       const vid_t *column_offset, const vid_t *row_indices, const vid_t *destination, const vid_t *source,
                       const vid_t *in_degree, const vid_t *out_degree, vid_t n_dst, vid_t n_src, vid_t feature_size) {
  memset(dx, 0, sizeof(float) * (size_t)n_src * feature_size);
  for (vid_t d = 0; d < n_dst; d++) {
    const float *in = dy + (size_t)d * feature_size;
    for (vid_t e = column_offset[d]; e < column_offset[d + 1]; e++) {
      vid_t ls = row_indices[e];
      float w = weight ? weight[e] : norm_degree(in_degree, out_degree, source[ls], destination[d]);
      float *o = dx + (size_t)ls * feature_size;
      for (vid_t k = 0; k < feature_size; k++) { float t = in[k] * w; o[k] = o[k] + t; }
    }
  }
}

/* Same backward expressed over the CSR with e_w_b (what Gather_By_Src_From_Dst computes,
 * cuda/ntsCUDAFuseKernel.cuh:494-531 / cuda/ntsCUDAGraphOP.cu:901-1042). */
void orc_aggregate_bwd_csr(const float *dy, float *dx, const float *weight_b, const vid_t *row_offset,
                           const vid_t *column_indices, vid_t n_src, vid_t feature_size) {
  memset(dx, 0, sizeof(float) * (size_t)n_src * feature_size);
  for (vid_t s = 0; s < n_src; s++) {
    float *o = dx + (size_t)s * feature_size;
    for (vid_t j = row_offset[s]; j < row_offset[s + 1]; j++) {
      const float *in = dy + (size_t)column_indices[j] * feature_size;
      float w = weight_b ? weight_b[j] : 1.0f;
      for (vid_t k = 0; k < feature_size; k++) { float t = in[k] * w; o[k] = o[k] + t; }
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * GAT edge ops (the reference's GPU kernels are the definition; the CPU push-down ops
 * core/ntsPushdownGraphOp.hpp:749-960 compute the same forward).
 *
 * scatter_src_dst_to_msg_map, cuda/ntsCUDADistKernel.cuh:174-196:
 *   msg[e, 0:F] = X[row_indices[e], :],  msg[e, F:2F] = X[dst_local_id[d], :]. */
void orc_scatter_src_dst(const float *x, float *msg /*[E,2F]*/, const vid_t *column_offset, const vid_t *row_indices,
                         const vid_t *dst_local_id, vid_t n_dst, vid_t F) {
  for (vid_t d = 0; d < n_dst; d++)
    for (vid_t e = column_offset[d]; e < column_offset[d + 1]; e++) {
      memcpy(msg + (size_t)e * 2 * F, x + (size_t)row_indices[e] * F, sizeof(float) * F);
      memcpy(msg + (size_t)e * 2 * F + F, x + (size_t)dst_local_id[d] * F, sizeof(float) * F);
    }
}
/* gather_msg_to_src_dst_map, ibid. :81-99: backward of the above (sums both halves into dX). */
void orc_gather_src_dst(float *dx /*[S,F]*/, const float *dmsg /*[E,2F]*/, const vid_t *column_offset,
                        const vid_t *row_indices, const vid_t *dst_local_id, vid_t n_dst, vid_t n_src, vid_t F) {
  memset(dx, 0, sizeof(float) * (size_t)n_src * F);
  for (vid_t d = 0; d < n_dst; d++)
    for (vid_t e = column_offset[d]; e < column_offset[d + 1]; e++)
      for (vid_t k = 0; k < F; k++) {
        dx[(size_t)row_indices[e] * F + k] += dmsg[(size_t)e * 2 * F + k];
        dx[(size_t)dst_local_id[d] * F + k] += dmsg[(size_t)e * 2 * F + F + k];
      }
}
/* get_node_max + edge_softmax_forward_norm_block, ibid. :371-388, :318-368 (feature_size 1):
 *   a[e] = exp(m[e] - max_d) / sum_{e' in col d} exp(m[e'] - max_d). */
void orc_edge_softmax_fwd(const float *m, float *a, const vid_t *column_offset, vid_t n_dst) {
  for (vid_t d = 0; d < n_dst; d++) {
    vid_t s = column_offset[d], t = column_offset[d + 1];
    if (s == t) continue;
    float mx = m[s];
    for (vid_t e = s + 1; e < t; e++) mx = m[e] > mx ? m[e] : mx;
    float sum = 0.0f;
    for (vid_t e = s; e < t; e++) sum += expf(m[e] - mx);
    for (vid_t e = s; e < t; e++) a[e] = expf(m[e] - mx) / sum;
  }
}
/* edge_softmax_backward_block, ibid. :440-484: dm[e] = da[e]*a[e] - a[e]*sum_{e'} da[e']*a[e']. */
void orc_edge_softmax_bwd(const float *da, const float *a, float *dm, const vid_t *column_offset, vid_t n_dst) {
  for (vid_t d = 0; d < n_dst; d++) {
    vid_t s = column_offset[d], t = column_offset[d + 1];
    float agg = 0.0f;
    for (vid_t e = s; e < t; e++) agg += da[e] * a[e];
    for (vid_t e = s; e < t; e++) dm[e] = da[e] * a[e] - agg * a[e];
  }
}
/* gather_msg_to_dst, ibid. :218-232: Y[d,:] = sum_{e in col d} msg[e,:]. */
void orc_gather_msg_to_dst(float *y /*[V,F]*/, const float *msg /*[E,F]*/, const vid_t *column_offset, vid_t n_dst, vid_t F) {
  memset(y, 0, sizeof(float) * (size_t)n_dst * F);
  for (vid_t d = 0; d < n_dst; d++)
    for (vid_t e = column_offset[d]; e < column_offset[d + 1]; e++)
      for (vid_t k = 0; k < F; k++) y[(size_t)d * F + k] += msg[(size_t)e * F + k];
}
/* scatter_dst_to_msg, ibid. :119-133: msg[e,:] = Y[d,:] (backward of the above). */
void orc_scatter_dst_to_msg(float *msg /*[E,F]*/, const float *y /*[V,F]*/, const vid_t *column_offset, vid_t n_dst, vid_t F) {
  for (vid_t d = 0; d < n_dst; d++)
    for (vid_t e = column_offset[d]; e < column_offset[d + 1]; e++)
      memcpy(msg + (size_t)e * F, y + (size_t)d * F, sizeof(float) * F);
}

/* The GAT layer as the toolkit composes it, toolkits/GAT_SAMPLE_ALL_MULTI.hpp:383-464, for
 * H = X*W already applied and the edge NN a = W_att[2F,1] (Parameter::forward is a bias-free
 * matmul): m = leaky_relu([h_src, h_dst] . a, 0.2); alpha = edge_softmax(m);
 * out[d,:] = sum_e alpha[e] * h_src(e). Writes alpha[E], score_pre[E] (pre-activation) and out. */
void orc_gat_layer_fwd(const float *h /*[S,F]*/, const float *att /*[2F]*/, const vid_t *column_offset,
                       const vid_t *row_indices, const vid_t *dst_local_id, vid_t n_dst, vid_t F,
                       float *score_pre /*[E]*/, float *alpha /*[E]*/, float *out /*[V,F]*/) {
  vid_t E = column_offset[n_dst];
  float *m = (float *)malloc(sizeof(float) * (E ? E : 1));
  for (vid_t d = 0; d < n_dst; d++)
    for (vid_t e = column_offset[d]; e < column_offset[d + 1]; e++) {
      float s = 0.0f;
      const float *hs = h + (size_t)row_indices[e] * F, *hd = h + (size_t)dst_local_id[d] * F;
      for (vid_t k = 0; k < F; k++) s += hs[k] * att[k];
      for (vid_t k = 0; k < F; k++) s += hd[k] * att[F + k];
      score_pre[e] = s;
      m[e] = s > 0.0f ? s : 0.2f * s;
    }
  orc_edge_softmax_fwd(m, alpha, column_offset, n_dst);
  memset(out, 0, sizeof(float) * (size_t)n_dst * F);
  for (vid_t d = 0; d < n_dst; d++)
    for (vid_t e = column_offset[d]; e < column_offset[d + 1]; e++)
      for (vid_t k = 0; k < F; k++) out[(size_t)d * F + k] += alpha[e] * h[(size_t)row_indices[e] * F + k];
  free(m);
}

/* Backward of the composed layer (chain rule through the same five ops, accumulating in
 * double so it can serve as the 1e-5 reference for the fused kernel). */
void orc_gat_layer_bwd(const float *h, const float *att, const float *dout /*[V,F]*/, const float *score_pre,
                       const float *alpha, const vid_t *column_offset, const vid_t *row_indices,
                       const vid_t *dst_local_id, vid_t n_dst, vid_t n_src, vid_t F, float *dh /*[S,F]*/, float *datt /*[2F]*/) {
  double *dH = (double *)calloc((size_t)n_src * F, sizeof(double));
  double *dA = (double *)calloc((size_t)2 * F, sizeof(double));
  for (vid_t d = 0; d < n_dst; d++) {
    vid_t s = column_offset[d], t = column_offset[d + 1];
    const float *go = dout + (size_t)d * F;
    const float *hd = h + (size_t)dst_local_id[d] * F;
    double agg = 0.0;
    for (vid_t e = s; e < t; e++) {
      const float *hs = h + (size_t)row_indices[e] * F;
      double da = 0.0;
      for (vid_t k = 0; k < F; k++) da += (double)go[k] * hs[k];
      agg += da * alpha[e];
    }
    for (vid_t e = s; e < t; e++) {
      const float *hs = h + (size_t)row_indices[e] * F;
      double da = 0.0;
      for (vid_t k = 0; k < F; k++) da += (double)go[k] * hs[k];
      double dm = alpha[e] * (da - agg);
      double ds = score_pre[e] > 0.0f ? dm : 0.2 * dm;
      double *ghs = dH + (size_t)row_indices[e] * F, *ghd = dH + (size_t)dst_local_id[d] * F;
      for (vid_t k = 0; k < F; k++) {
        ghs[k] += (double)alpha[e] * go[k] + ds * att[k];
        ghd[k] += ds * att[F + k];
        dA[k] += ds * hs[k];
        dA[F + k] += ds * hd[k];
      }
    }
  }
  for (size_t i = 0; i < (size_t)n_src * F; i++) dh[i] = (float)dH[i];
  for (vid_t k = 0; k < 2 * F; k++) datt[k] = (float)dA[k];
  free(dH);
  free(dA);
}

/* ------------------------------------------------------------------------------------------
 * Hotness-aware cache index. core/ntsDataloador.hpp:440-452 (set_cache_index):
 * cache_map[id] = super_batch_id, cache_location[id] = position in the id list. */
void orc_set_cache_index(vid_t *cache_map, vid_t *cache_location, vid_t super_batch_id, const vid_t *cache_ids, vid_t n) {
  for (vid_t i = 0; i < n; i++) { cache_map[cache_ids[i]] = super_batch_id; cache_location[cache_ids[i]] = i; }
}

/* ------------------------------------------------------------------------------------------
 * Hotness pre-sampling. core/ntsBaseOp.hpp:333-399 (get_most_neighbor, the overload preSample
 * :415-541 calls): counts start as the indicator of the super-batch's seeds and are pushed
 * `layers-1` times along the in-edges of the FULL graph (new[src] += old[dst] for every
 * in-neighbour src of every dst with old[dst] > 0); the counts are sorted descending,
 * total_sample_num = (index of the first zero) + 1, cache_num = (u32)(total_sample_num * cache_rate)
 * in float arithmetic, pivot = sorted[cache_num]; the ids with count >= pivot are collected, at most
 * cache_num of them. The reference collects them from an OpenMP loop (arrival order); the serial,
 * ascending-id order is restated (it is what the shipped .bin and the 1-thread run contain).
 * Returns cache_num; writes the final counts to counts_out[V] if not NULL. */
static int cmp_desc_u32(const void *a, const void *b) {
  vid_t x = *(const vid_t *)a, y = *(const vid_t *)b;
  return x < y ? 1 : (x > y ? -1 : 0);
}
vid_t orc_hotness(const vid_t *seeds, vid_t n_seeds, const vid_t *g_column_offset, const vid_t *g_row_indices, vid_t n_vertices,
                  int layers, float cache_rate, vid_t *cache_ids /*[n_vertices]*/, vid_t *counts_out) {
  vid_t *oldc = (vid_t *)calloc(n_vertices, sizeof(vid_t)), *newc = (vid_t *)calloc(n_vertices, sizeof(vid_t));
  for (vid_t i = 0; i < n_seeds; i++) oldc[seeds[i]] = 1;
  for (int layer = 1; layer < layers; layer++) {
    if (layer != 1) { vid_t *t = oldc; oldc = newc; newc = t; memset(newc, 0, sizeof(vid_t) * n_vertices); }
    for (vid_t i = 0; i < n_vertices; i++)
      if (oldc[i] > 0)
        for (vid_t e = g_column_offset[i]; e < g_column_offset[i + 1]; e++) newc[g_row_indices[e]] += oldc[i];
  }
  if (counts_out) memcpy(counts_out, newc, sizeof(vid_t) * n_vertices);
  memcpy(oldc, newc, sizeof(vid_t) * n_vertices);
  qsort(oldc, n_vertices, sizeof(vid_t), cmp_desc_u32);
  vid_t total = n_vertices; /* the reference leaves it uninitialised when no count is zero */
  for (vid_t i = 0; i < n_vertices; i++) if (oldc[i] == 0) { total = i + 1; break; }
  vid_t cache_num = (vid_t)((float)total * cache_rate);
  if (cache_num >= n_vertices) cache_num = n_vertices - 1;
  vid_t pivot = oldc[cache_num], idx = 0;
  for (vid_t i = 0; i < n_vertices && idx < cache_num; i++)
    if (newc[i] >= pivot) cache_ids[idx++] = i;
  free(oldc);
  free(newc);
  return cache_num;
}
