// aggregate.cu -- sparse aggregation forward / backward as segment reductions.
//
// Replaces (reference file:line)
//   Gather_By_Dst_From_Src_Spmm  (cuSPARSE SpMM, CSC^T * X)   cuda/ntsCUDAGraphOP.cu:425-587
//   Gather_By_Src_From_Dst_Spmm  (cuSPARSE SpMM, CSR * dY)     cuda/ntsCUDAGraphOP.cu:901-1042
//   Push_From_Dst_To_Src_Spmm    (cuSPARSE SpMM, CSC * dY)     cuda/ntsCUDAGraphOP.cu:621-770
//   and the legacy atomics kernels aggregate_kernel_from_src_with_weight / push_kernel_from_dst_with_weight /
//   aggregate_kernel_from_dst_with_weight  cuda/ntsCUDAFuseKernel.cuh:272-353, 494-531
// Semantic definition = the CPU op MiniBatchFuseOp (core/ntsMiniBatchGraphOp.hpp:153-182, 214-268):
//   out[r,:] = sum over the segment of r, in stored order, of w[j] * in[idx[j],:], multiply then add.
//
// HBM bound (<= 0.5 flop/byte); tensor cores are not used. Algorithmic bytes per launch:
//   E*(4 idx + 4 w + 4F row) + (R+1)*4 offsets + R*4F output.
// Mapping: one warp per output row; the row's accumulators live in registers (CHUNK vectors per
// lane, so a 602-float row is 10 float2 per lane); the segment's (idx,w) pairs are fetched 32 at
// a time, one per lane, and broadcast by shuffle; input rows are read with 64-/128-bit read-only
// streaming loads, two segment entries in flight per lane. No atomics, every output row is
// written exactly once (empty segment -> zeros), summation order == the CPU op's.
#include "common.cuh"

constexpr int AGG_THREADS = 256;
// 7 resident blocks per SM, not 8: the free 256-thread slot lets the next batch's sampling kernels and the gradient exchange start
// under a running aggregation instead of waiting for its tail (same-call sweep, profiles/r2_sweep_occupancy.txt: 0.1554 -> 0.1522 ms per step)
static int g_agg_blocks_per_sm = 7, g_agg_persistent = 1, g_agg_long_rows = 0, g_agg_pipe_wide = 1, g_agg_short_rows = 0, g_agg_deep_small = 1;  // pipe: 0 never, 1 rows of <= 64 vectors, 2 always
void nb_agg_set_option(int which, int value) {
  if (which == 0) g_agg_blocks_per_sm = value < 1 ? 1 : value > 8 ? 8 : value;
  else if (which == 1) g_agg_persistent = value;
  else if (which == 2) g_agg_long_rows = value;
  else if (which == 4) g_agg_short_rows = value;
  else if (which == 5) g_agg_deep_small = value;
  else g_agg_pipe_wide = value;
}

// UNR = segment entries whose row loads are issued before the first accumulate (UNR*CHUNK independent vector loads per
// lane in flight); narrow rows (CHUNK 1-2, e.g. F=128) take 4 entries at a time, wide rows 2. Accumulation stays in
// stored order, so the result does not depend on UNR.
// Optional rank-2 epilogue (used by the fused GAT backward): out[r,:] += e1[r] * va[:] + e2[r] * vb[:].
struct SegEpilogue {
  const float *e1, *e2;  // per output row scalars (e2 may be NULL)
  const float *va, *vb;  // feature-length vectors
  // fused GAT backward (gat.cu): the row's weights and scalars are derived on the fly from CSC-ordered per-edge arrays instead of
  // being staged by a separate kernel: for CSR entry j, e = c2c[j]: w = alpha[e]; e1(r) = sum_j ds[e]; e2(r) = dsum[src_to_dst[r]]
  // (0 when the source is not a dst). e1 / e2 are also written to rs_out / dd_out for the attention-gradient pass.
  const uint32_t *c2c = nullptr, *src_to_dst = nullptr;
  const float *alpha = nullptr, *ds = nullptr, *dsum = nullptr;
  float *rs_out = nullptr, *dd_out = nullptr;
};

// GATHERED (the bottom hop fused with the feature gather): `in` is the feature table and idx[j] a PACKED gather index of the
// sampler (nb_layer_view::gather_index): bits 0-30 = the global id of the edge's source, bit 31 = "this batch reads that row more
// than once". X0 = table[source] is never written; rows with the bit set are loaded with an L2 evict_last hint, the others with
// evict_first, so the re-reads (41 % of the row reads on the Reddit shape) hit L2 instead of being flushed by the single-use
// rows in between. No extra load per edge: the hint travels in the index.
template <int VEC, int CHUNK, int UNR, bool GATHERED>
__global__ void __launch_bounds__(AGG_THREADS)
k_segment_reduce(const float *__restrict__ in, float *__restrict__ out, const float *__restrict__ weight,
                 const uint32_t *__restrict__ idx, const uint32_t *__restrict__ offsets, uint32_t n_rows,
                 const uint32_t *__restrict__ n_rows_dev, uint32_t nvec, uint64_t pitch, uint64_t out_pitch,
                 SegEpilogue epi = SegEpilogue{nullptr, nullptr, nullptr, nullptr}) {
  const unsigned lane = lane_id();
  const unsigned warp = (blockIdx.x * AGG_THREADS + threadIdx.x) >> 5;
  const unsigned warps = (gridDim.x * AGG_THREADS) >> 5;
  if (n_rows_dev) n_rows = min(n_rows, *n_rows_dev);
  uint64_t pol_keep = 0, pol_once = 0;
  if (GATHERED) { pol_keep = l2_policy_evict_last(); pol_once = l2_policy_evict_first(); }
  for (unsigned r = warp; r < n_rows; r += warps) {
    const uint32_t beg = offsets[r], end = offsets[r + 1];
    float gat_aux = 0.f, gat_s1 = 0.f, gat_s2 = 0.f;
    for (unsigned c0 = 0; c0 < nvec; c0 += 32 * CHUNK) {  // one pass unless the row is wider than 32*CHUNK vectors
      Vec<VEC> acc[CHUNK];
#pragma unroll
      for (int c = 0; c < CHUNK; c++) acc[c].zero();
      for (uint32_t j0 = beg; j0 < end; j0 += 32) {
        const uint32_t cnt = min(32u, end - j0);
        uint32_t my_idx = 0, my_keep = 0;
        float my_w = 1.0f;
        if (lane < cnt) {
          my_idx = idx[j0 + lane];
          if (epi.c2c) {
            const uint32_t e = epi.c2c[j0 + lane];
            my_w = epi.alpha[e];
            if (c0 == 0) gat_aux += epi.ds[e];
          } else if (weight) my_w = weight[j0 + lane];
          if (GATHERED) { my_keep = my_idx >> 31; my_idx &= 0x7fffffffu; }
        }
        for (uint32_t t = 0; t < cnt; t += UNR) {
          Vec<VEC> x[UNR][CHUNK];
          float w[UNR];
#pragma unroll
          for (int u = 0; u < UNR; u++) {
            // entries past the end of the batch re-read entry t (cheap, cached) and are not accumulated
            const uint32_t tt = t + u < cnt ? t + u : t;
            const uint32_t s = __shfl_sync(FULL_MASK, my_idx, tt);
            w[u] = __shfl_sync(FULL_MASK, my_w, tt);
            const float *p = in + (uint64_t)s * pitch;
            uint64_t pol = 0;
            if (GATHERED) pol = __shfl_sync(FULL_MASK, my_keep, tt) ? pol_keep : pol_once;
            if (t + u < cnt) {
#pragma unroll
              for (int c = 0; c < CHUNK; c++) {
                const unsigned k = c0 + c * 32 + lane;
                if (k < nvec) {
                  if (GATHERED) x[u][c].load_hint(p + (uint64_t)k * VEC, pol);
                  else x[u][c].load(p + (uint64_t)k * VEC);
                }
              }
            }
          }
#pragma unroll
          for (int u = 0; u < UNR; u++) {
            if (t + u < cnt) {
#pragma unroll
              for (int c = 0; c < CHUNK; c++) {
                const unsigned k = c0 + c * 32 + lane;
                if (k < nvec) acc[c].axpy(x[u][c], w[u]);
              }
            }
          }
        }
      }
      if (epi.c2c && c0 == 0) {
        gat_s1 = warp_sum(gat_aux);
        const uint32_t d = epi.src_to_dst[r];
        gat_s2 = d != 0xffffffffu ? epi.dsum[d] : 0.f;
        if (lane == 0) { epi.rs_out[r] = gat_s1; epi.dd_out[r] = gat_s2; }
      }
      if (epi.e1 || epi.c2c) {
        const float s1 = epi.c2c ? gat_s1 : epi.e1[r], s2 = epi.c2c ? gat_s2 : (epi.e2 ? epi.e2[r] : 0.f);
        const bool two = epi.c2c || epi.e2;
#pragma unroll
        for (int c = 0; c < CHUNK; c++) {
          const unsigned k = c0 + c * 32 + lane;
          if (k < nvec) {
            Vec<VEC> a, b;
            // va / vb are the same few lines for every row of the launch: read them through L1. As streaming (L1-bypassing)
            // loads they all land on the same L2 slices and serialise there (150K rows x 1 KB: ~35 us of the GAT backward)
            a.load_cached(epi.va + (uint64_t)k * VEC);
            acc[c].axpy(a, s1);
            if (two) { b.load_cached(epi.vb + (uint64_t)k * VEC); acc[c].axpy(b, s2); }
          }
        }
      }
      float *o = out + (uint64_t)r * out_pitch;
#pragma unroll
      for (int c = 0; c < CHUNK; c++) {
        const unsigned k = c0 + c * 32 + lane;
        if (k < nvec) acc[c].store(o + (uint64_t)k * VEC);
      }
    }
  }
}

// ---- load-balanced variant (opt-in: nb_set_option("agg_long_rows", 1)) --------------------------------------------------------
// Measured side by side on the B200 (same box, same call) this variant costs ~2 us per launch over the plain kernel above on the
// headline workload, whose sampled segments never exceed the fanout -- so it is not the default. It is the kernel to switch on
// for take-all layers or hub-heavy CSR backwards.
// Load balance. A warp walks its segment sequentially (that is what keeps the CPU summation order), so one hub row -- a source
// that thousands of sampled columns point at in the CSR backward, or a take-all column of a hub in the forward -- would keep a
// single warp busy long after the rest of the grid has finished (4 row loads in flight: ~3 GB/s per warp, ~0.15 us per entry).
// Segments longer than SEG_LONG are therefore only *recorded* (per block, in shared memory) by the warp that meets them and are
// reduced after the warp loop by the whole block: all 256 threads stage a batch of input rows in shared memory (up to 256
// independent vector loads in flight instead of 4), then every thread adds its own feature columns entry by entry -- the same
// order, the same rounding, the same bits as the warp path. No second launch, no global work list.
//
// Latency. Short segments (CSR backward: ~1.6 entries per row; top hops) are bound by the dependent chain
// offsets -> (idx, w) -> input rows -> store (~3.3 us per row measured), not by bandwidth. The warp loop is software pipelined:
// the offsets of the row two iterations ahead and the first 32 (idx, w) of the next row are requested before the current row
// is reduced, which leaves rows -> store on the critical path.
constexpr uint32_t SEG_LONG = 96;          // entries
constexpr uint32_t SEG_LOCAL_CAP = 32;     // long rows a block can queue; further ones are reduced by their warp (slow, correct)
constexpr uint32_t SEG_LONG_MAX_F = 2048;  // 8 accumulators per thread
constexpr uint32_t SEG_STAGE_BYTES = 16 * 1024;

template <int VEC>
__device__ __noinline__ void segment_block_reduce(uint32_t r, const float *__restrict__ in, float *__restrict__ out,
                                                     const float *__restrict__ weight, const uint32_t *__restrict__ idx,
                                                     const uint32_t *__restrict__ offsets, uint32_t nvec, uint64_t pitch, uint64_t out_pitch,
                                                     const SegEpilogue &epi, float *s_rows, uint32_t batch) {
  const uint32_t Fp = nvec * VEC, t = threadIdx.x;
  uint32_t *s_idx = (uint32_t *)(s_rows + (size_t)batch * Fp);
  float *s_w = (float *)(s_idx + batch);
  const uint32_t beg = offsets[r], end = offsets[r + 1];
  float acc[SEG_LONG_MAX_F / AGG_THREADS];
#pragma unroll
  for (int q = 0; q < (int)(SEG_LONG_MAX_F / AGG_THREADS); q++) acc[q] = 0.f;
  for (uint32_t j0 = beg; j0 < end; j0 += batch) {
    const uint32_t cnt = min(batch, end - j0);
    if (t < cnt) {
      s_idx[t] = idx[j0 + t];
      s_w[t] = weight ? weight[j0 + t] : 1.0f;
    }
    __syncthreads();
    for (uint32_t k = t; k < cnt * nvec; k += AGG_THREADS) {
      const uint32_t e = k / nvec, c = k - e * nvec;
      Vec<VEC> x;
      x.load(in + (uint64_t)s_idx[e] * pitch + (uint64_t)c * VEC);
      *reinterpret_cast<decltype(x.v) *>(s_rows + (size_t)e * Fp + (size_t)c * VEC) = x.v;
    }
    __syncthreads();
    for (uint32_t e = 0; e < cnt; e++) {
      const float w = s_w[e];
#pragma unroll
      for (int q = 0; q < (int)(SEG_LONG_MAX_F / AGG_THREADS); q++) {
        const uint32_t f = t + q * AGG_THREADS;
        if (f < Fp) acc[q] = __fadd_rn(acc[q], __fmul_rn(s_rows[(size_t)e * Fp + f], w));
      }
    }
    __syncthreads();
  }
  float s1 = 0.f, s2 = 0.f;
  if (epi.e1) { s1 = epi.e1[r]; s2 = epi.e2 ? epi.e2[r] : 0.f; }
#pragma unroll
  for (int q = 0; q < (int)(SEG_LONG_MAX_F / AGG_THREADS); q++) {
    const uint32_t f = t + q * AGG_THREADS;
    if (f < Fp) {
      float a = acc[q];
      if (epi.e1) {
        a = __fadd_rn(a, __fmul_rn(epi.va[f], s1));
        if (epi.e2) a = __fadd_rn(a, __fmul_rn(epi.vb[f], s2));
      }
      out[(uint64_t)r * out_pitch + f] = a;
    }
  }
}

// ("agg_short_rows", off by default: measured, it does not beat the warp-per-row kernel -- 13.0 vs 12.5 us on the top hop's 24K rows,
// 162 vs 133 us inside the GAT backward's 150K rows, where its 88-100 registers halve the occupancy; profiles/r2_gat_short_rows_ab.txt)
// Short rows -- the CSR of a sampled layer (backward: ~1-2 entries per source row) with rows of <= 32 vectors: one warp per row
// chains three dependent loads (offsets -> index/weight -> data) for a single 512-byte row, and the launch is bound by that
// latency times the number of waves. Here a warp owns ROWS consecutive rows: one load fetches their ROWS+1 offsets, one
// coalesced load the indices / weights of all their entries (they are contiguous), and the rows' data loads are in flight
// together. Per row the entries are still accumulated in stored order (mul, then add): same bits as k_segment_reduce.
template <int VEC, int ROWS, bool GAT>
__global__ void __launch_bounds__(AGG_THREADS)
k_segment_reduce_short(const float *__restrict__ in, float *__restrict__ out, const float *__restrict__ weight,
                       const uint32_t *__restrict__ idx, const uint32_t *__restrict__ offsets, uint32_t n_rows,
                       const uint32_t *__restrict__ n_rows_dev, uint32_t nvec, uint64_t pitch, uint64_t out_pitch, SegEpilogue epi) {
  constexpr uint32_t JOINT = 4;   // entries per row handled in the joint phase; longer rows finish one at a time, 4 loads in flight
  const unsigned lane = lane_id();
  const unsigned warp = (blockIdx.x * AGG_THREADS + threadIdx.x) >> 5;
  const unsigned warps = (gridDim.x * AGG_THREADS) >> 5;
  if (n_rows_dev) n_rows = min(n_rows, *n_rows_dev);
  const bool active = lane < nvec;
  const uint64_t col = (uint64_t)lane * VEC;
  Vec<VEC> va, vb;   // GAT: the two epilogue vectors, the same for every row
  if (GAT) { va.zero(); vb.zero(); if (active) { va.load_cached(epi.va + col); vb.load_cached(epi.vb + col); } }
  for (unsigned r0 = warp * ROWS; r0 < n_rows; r0 += warps * ROWS) {
    const unsigned nr = min((unsigned)ROWS, n_rows - r0);
    const uint32_t off = lane <= nr ? offsets[r0 + lane] : 0u;
    uint32_t beg[ROWS], len[ROWS];
    uint32_t maxlen = 0;
#pragma unroll
    for (int i = 0; i < ROWS; i++) {
      const uint32_t b = __shfl_sync(FULL_MASK, off, i), e = __shfl_sync(FULL_MASK, off, i + 1);
      beg[i] = b;
      len[i] = (unsigned)i < nr ? e - b : 0u;
      maxlen = max(maxlen, len[i]);
    }
    const uint32_t e0 = beg[0], total = __shfl_sync(FULL_MASK, off, nr) - e0;
    uint32_t my_idx = 0;
    float my_w = 1.0f, my_ds = 0.f;
    if (lane < total) {   // the first 32 entries of the group in one coalesced load
      my_idx = idx[e0 + lane];
      if (GAT) { const uint32_t e = epi.c2c[e0 + lane]; my_w = epi.alpha[e]; my_ds = epi.ds[e]; }
      else if (weight) my_w = weight[e0 + lane];
    }
    Vec<VEC> acc[ROWS];
    float s1[ROWS];
#pragma unroll
    for (int i = 0; i < ROWS; i++) { acc[i].zero(); s1[i] = 0.f; }
    const uint32_t joint = min(maxlen, JOINT);
    for (uint32_t j = 0; j < joint; j++) {
      Vec<VEC> x[ROWS];
      float w[ROWS];
#pragma unroll
      for (int i = 0; i < ROWS; i++) {
        if (j < len[i]) {   // warp-uniform
          const uint32_t e = beg[i] + j - e0;
          uint32_t s;
          if (e < 32) {
            s = __shfl_sync(FULL_MASK, my_idx, e); w[i] = __shfl_sync(FULL_MASK, my_w, e);
            if (GAT) s1[i] += __shfl_sync(FULL_MASK, my_ds, e);
          } else {
            s = idx[e0 + e];
            if (GAT) { const uint32_t ce = epi.c2c[e0 + e]; w[i] = epi.alpha[ce]; s1[i] += epi.ds[ce]; }
            else w[i] = weight ? weight[e0 + e] : 1.0f;
          }
          if (active) x[i].load(in + (uint64_t)s * pitch + col);
        }
      }
#pragma unroll
      for (int i = 0; i < ROWS; i++)
        if (j < len[i] && active) acc[i].axpy(x[i], w[i]);
    }
    if (maxlen > JOINT) {
#pragma unroll
      for (int i = 0; i < ROWS; i++) {
        for (uint32_t j = JOINT; j < len[i]; j += 4) {
          Vec<VEC> x[4];
          float w[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            if (j + u < len[i]) {
              const uint32_t e = beg[i] + j + u;
              const uint32_t s = idx[e];
              if (GAT) { const uint32_t ce = epi.c2c[e]; w[u] = epi.alpha[ce]; s1[i] += epi.ds[ce]; }
              else w[u] = weight ? weight[e] : 1.0f;
              if (active) x[u].load(in + (uint64_t)s * pitch + col);
            }
          }
#pragma unroll
          for (int u = 0; u < 4; u++)
            if (j + u < len[i] && active) acc[i].axpy(x[u], w[u]);
        }
      }
    }
    if (GAT) {   // out[r,:] += s1 va + s2 vb, s1 = sum of the row's ds, s2 = dsum of the dst this source also is (gat.cu)
      uint32_t d = 0xffffffffu;
      if (lane < nr) d = epi.src_to_dst[r0 + lane];
      float s2l = d != 0xffffffffu ? epi.dsum[d] : 0.f;
#pragma unroll
      for (int i = 0; i < ROWS; i++) {
        const float s2 = __shfl_sync(FULL_MASK, s2l, i);
        if ((unsigned)i < nr) {
          if (lane == 0) { epi.rs_out[r0 + i] = s1[i]; epi.dd_out[r0 + i] = s2; }
          if (active) { acc[i].axpy(va, s1[i]); acc[i].axpy(vb, s2); }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < ROWS; i++)
      if ((unsigned)i < nr && active) acc[i].store(out + (uint64_t)(r0 + i) * out_pitch + col);
  }
}

template <int VEC, int CHUNK, int UNR, bool PIPE>
__global__ void __launch_bounds__(AGG_THREADS)
k_segment_reduce_lb(const float *__restrict__ in, float *__restrict__ out, const float *__restrict__ weight,
                 const uint32_t *__restrict__ idx, const uint32_t *__restrict__ offsets, uint32_t n_rows,
                 const uint32_t *__restrict__ n_rows_dev, uint32_t nvec, uint64_t pitch, uint64_t out_pitch,
                 SegEpilogue epi = SegEpilogue{nullptr, nullptr, nullptr, nullptr}, uint32_t long_batch = 0) {
  extern __shared__ __align__(16) float s_stage[];  // long_batch staged rows + their (idx, w); empty when long_batch == 0
  __shared__ uint32_t s_long[SEG_LOCAL_CAP];
  __shared__ uint32_t s_nlong;
  if (long_batch) {
    if (threadIdx.x == 0) s_nlong = 0;
    __syncthreads();
  }
  const unsigned lane = lane_id();
  const unsigned warp = (blockIdx.x * AGG_THREADS + threadIdx.x) >> 5;
  const unsigned warps = (gridDim.x * AGG_THREADS) >> 5;
  if (n_rows_dev) n_rows = min(n_rows, *n_rows_dev);
  // pipeline registers: current row (beg, end, first 32 entries), next row's offsets
  uint32_t beg = 0, end = 0, nbeg = 0, nend = 0, cur_idx = 0;
  float cur_w = 1.0f, cur_s1 = 0.f, cur_s2 = 0.f;
  if (PIPE && epi.e1 && warp < n_rows) { cur_s1 = epi.e1[warp]; cur_s2 = epi.e2 ? epi.e2[warp] : 0.f; }
  if (PIPE) {
    if (warp < n_rows) { beg = offsets[warp]; end = offsets[warp + 1]; }
    if (warp + warps < n_rows) { nbeg = offsets[warp + warps]; nend = offsets[warp + warps + 1]; }
    if (lane < end - beg) {
      cur_idx = idx[beg + lane];
      if (weight) cur_w = weight[beg + lane];
    }
  }
  for (unsigned r = warp; r < n_rows; r += warps) {
    uint32_t nnbeg = 0, nnend = 0, nxt_idx = 0;
    float nxt_w = 1.0f, nxt_s1 = 0.f, nxt_s2 = 0.f;
    if (PIPE) {
      if (epi.e1 && r + warps < n_rows) { nxt_s1 = epi.e1[r + warps]; nxt_s2 = epi.e2 ? epi.e2[r + warps] : 0.f; }
      if (r + 2 * warps < n_rows) { nnbeg = offsets[r + 2 * warps]; nnend = offsets[r + 2 * warps + 1]; }
      if (lane < nend - nbeg) {
        nxt_idx = idx[nbeg + lane];
        if (weight) nxt_w = weight[nbeg + lane];
      }
    } else {
      // wide rows are bandwidth bound: the extra requests and registers of the pipeline cost more than the latency they hide
      beg = offsets[r];
      end = offsets[r + 1];
      cur_idx = 0;
      cur_w = 1.0f;
      if (lane < end - beg) {
        cur_idx = idx[beg + lane];
        if (weight) cur_w = weight[beg + lane];
      }
      if (epi.e1) { cur_s1 = epi.e1[r]; cur_s2 = epi.e2 ? epi.e2[r] : 0.f; }
    }
    bool queued = false;
    if (long_batch && end - beg > SEG_LONG) {
      uint32_t k = 0;
      if (lane == 0) k = atomicAdd(&s_nlong, 1u);
      k = __shfl_sync(FULL_MASK, k, 0);
      if (k < SEG_LOCAL_CAP) {
        if (lane == 0) s_long[k] = r;
        queued = true;
      }
    }
    if (!queued) {
      for (unsigned c0 = 0; c0 < nvec; c0 += 32 * CHUNK) {  // one pass unless the row is wider than 32*CHUNK vectors
        Vec<VEC> acc[CHUNK];
#pragma unroll
        for (int c = 0; c < CHUNK; c++) acc[c].zero();
        for (uint32_t j0 = beg; j0 < end; j0 += 32) {
          const uint32_t cnt = min(32u, end - j0);
          uint32_t my_idx = cur_idx;
          float my_w = cur_w;
          if (j0 != beg) {
            my_idx = 0;
            my_w = 1.0f;
            if (lane < cnt) {
              my_idx = idx[j0 + lane];
              if (weight) my_w = weight[j0 + lane];
            }
          }
          for (uint32_t t = 0; t < cnt; t += UNR) {
            Vec<VEC> x[UNR][CHUNK];
            float w[UNR];
#pragma unroll
            for (int u = 0; u < UNR; u++) {
              // entries past the end of the batch re-read entry t (cheap, cached) and are not accumulated
              const uint32_t tt = t + u < cnt ? t + u : t;
              const uint32_t s = __shfl_sync(FULL_MASK, my_idx, tt);
              w[u] = __shfl_sync(FULL_MASK, my_w, tt);
              const float *p = in + (uint64_t)s * pitch;
              if (t + u < cnt) {
#pragma unroll
                for (int c = 0; c < CHUNK; c++) {
                  const unsigned k = c0 + c * 32 + lane;
                  if (k < nvec) x[u][c].load(p + (uint64_t)k * VEC);
                }
              }
            }
#pragma unroll
            for (int u = 0; u < UNR; u++) {
              if (t + u < cnt) {
#pragma unroll
                for (int c = 0; c < CHUNK; c++) {
                  const unsigned k = c0 + c * 32 + lane;
                  if (k < nvec) acc[c].axpy(x[u][c], w[u]);
                }
              }
            }
          }
        }
        if (epi.e1) {
          const float s1 = cur_s1, s2 = cur_s2;
#pragma unroll
          for (int c = 0; c < CHUNK; c++) {
            const unsigned k = c0 + c * 32 + lane;
            if (k < nvec) {
              // va / vb are the same few lines for every row of the launch: read them through L1. As streaming (L1-bypassing)
              // loads they all land on the same L2 slices and serialise there (150K rows x 1 KB: ~100 us of the GAT backward)
              Vec<VEC> a, b;
              a.load_cached(epi.va + (uint64_t)k * VEC);
              acc[c].axpy(a, s1);
              if (epi.e2) { b.load_cached(epi.vb + (uint64_t)k * VEC); acc[c].axpy(b, s2); }
            }
          }
        }
        float *o = out + (uint64_t)r * out_pitch;
#pragma unroll
        for (int c = 0; c < CHUNK; c++) {
          const unsigned k = c0 + c * 32 + lane;
          if (k < nvec) acc[c].store(o + (uint64_t)k * VEC);
        }
      }
    }
    if (PIPE) { beg = nbeg; end = nend; nbeg = nnbeg; nend = nnend; cur_idx = nxt_idx; cur_w = nxt_w; cur_s1 = nxt_s1; cur_s2 = nxt_s2; }
  }
  if (long_batch) {
    __syncthreads();
    const uint32_t nl = min(s_nlong, SEG_LOCAL_CAP);
    for (uint32_t i = 0; i < nl; i++)
      segment_block_reduce<VEC>(s_long[i], in, out, weight, idx, offsets, nvec, pitch, out_pitch, epi, s_stage, long_batch);
  }
}

// out[idx[e],:] += w[e] * in[d,:] for e in column d. One warp per dst row; the dY row is read once
// into registers and pushed with vector reductions (red.global.add.v2/v4.f32, sm_90+).
__device__ __forceinline__ void red_add(float *p, const Vec<4> &x, float w) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(x.v.x * w), "f"(x.v.y * w), "f"(x.v.z * w), "f"(x.v.w * w) : "memory");
}
__device__ __forceinline__ void red_add(float *p, const Vec<2> &x, float w) {
  asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(x.v.x * w), "f"(x.v.y * w) : "memory");
}
__device__ __forceinline__ void red_add(float *p, const Vec<1> &x, float w) { atomicAdd(p, x.v * w); }

template <int VEC, int CHUNK>
__global__ void __launch_bounds__(AGG_THREADS)
k_push(const float *__restrict__ in, float *__restrict__ out, const float *__restrict__ weight, const uint32_t *__restrict__ idx,
       const uint32_t *__restrict__ offsets, uint32_t n_rows, uint32_t nvec, uint64_t pitch) {
  const unsigned lane = lane_id();
  const unsigned warp = (blockIdx.x * AGG_THREADS + threadIdx.x) >> 5;
  const unsigned warps = (gridDim.x * AGG_THREADS) >> 5;
  for (unsigned r = warp; r < n_rows; r += warps) {
    const uint32_t beg = offsets[r], end = offsets[r + 1];
    if (beg == end) continue;
    const float *p = in + (uint64_t)r * pitch;
    for (unsigned c0 = 0; c0 < nvec; c0 += 32 * CHUNK) {
      Vec<VEC> x[CHUNK];
#pragma unroll
      for (int c = 0; c < CHUNK; c++) {
        const unsigned k = c0 + c * 32 + lane;
        if (k < nvec) x[c].load(p + (uint64_t)k * VEC);
      }
      for (uint32_t j = beg; j < end; j++) {
        const uint32_t s = idx[j];
        const float w = weight ? weight[j] : 1.0f;
        float *o = out + (uint64_t)s * pitch;
#pragma unroll
        for (int c = 0; c < CHUNK; c++) {
          const unsigned k = c0 + c * 32 + lane;
          if (k < nvec) red_add(o + (uint64_t)k * VEC, x[c], w);
        }
      }
    }
  }
}

template <int VEC>
static int launch_segment(nb_ctx *ctx, bool push, const float *in, float *out, const float *w, const uint32_t *idx,
                          const uint32_t *offsets, uint32_t n_rows, uint32_t F, const uint32_t *n_rows_dev, uint64_t in_pitch,
                          uint64_t out_pitch, SegEpilogue epi, bool packed_index, int shape = 0) {
  const uint32_t nvec = F / VEC;
  if (shape == NB_SEG_SHORT_ROWS && g_agg_short_rows && !push && !packed_index && !epi.e1 && nvec <= 32) {
    constexpr int ROWS = 4;
    const unsigned grid = nb_grid(n_rows, (AGG_THREADS / 32) * ROWS, 8);
    if (epi.c2c) k_segment_reduce_short<VEC, ROWS, true><<<grid, AGG_THREADS, 0, ctx->stream>>>(in, out, w, idx, offsets, n_rows, n_rows_dev, nvec, in_pitch, out_pitch, epi);
    else k_segment_reduce_short<VEC, ROWS, false><<<grid, AGG_THREADS, 0, ctx->stream>>>(in, out, w, idx, offsets, n_rows, n_rows_dev, nvec, in_pitch, out_pitch, epi);
    NB_LAUNCH_CHECK(ctx);
    return NB_OK;
  }
  // block path for long segments: staging batch sized to SEG_STAGE_BYTES of dynamic shared memory
  uint32_t long_batch = 0;
  size_t smem = 0;
  if (!push && !packed_index && !epi.c2c && g_agg_long_rows && F <= SEG_LONG_MAX_F) {
    long_batch = SEG_STAGE_BYTES / (F * 4u);
    if (long_batch > 64) long_batch = 64;
    if (long_batch < 2) long_batch = 2;
    smem = (size_t)long_batch * F * 4 + (size_t)long_batch * 8;
  }
  // persistent grid (blocks loop over rows) by default; "agg_persistent"=0 launches one warp per row so that blocks retire
  // every few microseconds and a concurrent higher-priority stream (the sampler of the next batch) gets SM slots in between
  const unsigned grid = g_agg_persistent ? nb_grid(n_rows, AGG_THREADS / 32, g_agg_blocks_per_sm)
                                         : (unsigned)((n_rows + AGG_THREADS / 32 - 1) / (AGG_THREADS / 32));
  const uint32_t per_lane = (nvec + 31) / 32;
#define NB_SEG(C)                                                                                              \
  do {                                                                                                         \
    if (push) k_push<VEC, C><<<grid, AGG_THREADS, 0, ctx->stream>>>(in, out, w, idx, offsets, n_rows, nvec, F); \
    else if (long_batch && (g_agg_pipe_wide == 2 || (C <= 2 && g_agg_pipe_wide == 1))) k_segment_reduce_lb<VEC, C, (C <= 2 ? 4 : C <= 4 ? 3 : 2), true><<<grid, AGG_THREADS, smem, ctx->stream>>>(in, out, w, idx, offsets, n_rows, n_rows_dev, nvec, in_pitch, out_pitch, epi, long_batch); \
    else if (long_batch) k_segment_reduce_lb<VEC, C, (C <= 2 ? 4 : C <= 4 ? 3 : 2), false><<<grid, AGG_THREADS, smem, ctx->stream>>>(in, out, w, idx, offsets, n_rows, n_rows_dev, nvec, in_pitch, out_pitch, epi, long_batch); \
    else if (packed_index) k_segment_reduce<VEC, C, (C <= 2 ? 4 : C <= 4 ? 3 : 2), true><<<grid, AGG_THREADS, 0, ctx->stream>>>(in, out, w, idx, offsets, n_rows, n_rows_dev, nvec, in_pitch, out_pitch, epi); \
    else k_segment_reduce<VEC, C, (C <= 2 ? 4 : C <= 4 ? 3 : 2), false><<<grid, AGG_THREADS, 0, ctx->stream>>>(in, out, w, idx, offsets, n_rows, n_rows_dev, nvec, in_pitch, out_pitch, epi); \
  } while (0)
  // a launch whose rows all fit on the GPU at once (a top hop: 1024 columns) is pure latency: the row's entries x load latency
  // / loads in flight. Registers are free there, so 16 (8) entries are in flight instead of 4.
  const bool small = g_agg_deep_small && !push && !packed_index && !long_batch && n_rows <= (unsigned)ctx->sm_count * 16u;
  if (small && per_lane <= 1) k_segment_reduce<VEC, 1, 16, false><<<grid, AGG_THREADS, 0, ctx->stream>>>(in, out, w, idx, offsets, n_rows, n_rows_dev, nvec, in_pitch, out_pitch, epi);
  else if (small && per_lane <= 2) k_segment_reduce<VEC, 2, 8, false><<<grid, AGG_THREADS, 0, ctx->stream>>>(in, out, w, idx, offsets, n_rows, n_rows_dev, nvec, in_pitch, out_pitch, epi);
  else if (per_lane <= 1) NB_SEG(1);
  else if (per_lane <= 2) NB_SEG(2);
  else if (per_lane <= 4) NB_SEG(4);
  else if (per_lane <= 5) NB_SEG(5);
  else if (per_lane <= 6) NB_SEG(6);
  else if (per_lane <= 8) NB_SEG(8);
  else if (per_lane <= 10) NB_SEG(10);
  else NB_SEG(12);
#undef NB_SEG
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
}

// CSR segment reduction of the fused GAT backward: dh[s,:] = sum_j alpha[c2c[j]] dout[col[j],:] + (sum_j ds[c2c[j]]) va + dsum[src_to_dst[s]] vb
int nb_run_segment_gat(nb_ctx *ctx, const float *dout, float *dh, const uint32_t *column_indices, const uint32_t *row_offset, uint32_t n_src,
                       uint32_t F, const uint32_t *c2c, const float *alpha, const float *ds, const float *dsum, const uint32_t *src_to_dst,
                       const float *va, const float *vb, float *rs_out, float *dd_out) {
  if (n_src == 0) return NB_OK;
  SegEpilogue epi{nullptr, nullptr, va, vb};
  epi.c2c = c2c; epi.alpha = alpha; epi.ds = ds; epi.dsum = dsum; epi.src_to_dst = src_to_dst; epi.rs_out = rs_out; epi.dd_out = dd_out;
  int vec = nb_pick_vec(F, dout, F, dh, F);
  if (vec > 1 && (((uintptr_t)va | (uintptr_t)vb) % (4 * vec))) vec = 1;
  if (vec == 4) return launch_segment<4>(ctx, false, dout, dh, nullptr, column_indices, row_offset, n_src, F, nullptr, F, F, epi, false, NB_SEG_SHORT_ROWS);
  if (vec == 2) return launch_segment<2>(ctx, false, dout, dh, nullptr, column_indices, row_offset, n_src, F, nullptr, F, F, epi, false, NB_SEG_SHORT_ROWS);
  return launch_segment<1>(ctx, false, dout, dh, nullptr, column_indices, row_offset, n_src, F, nullptr, F, F, epi, false, NB_SEG_SHORT_ROWS);
}

int nb_run_segment(nb_ctx *ctx, bool push, const float *in, float *out, const float *w, const uint32_t *idx,
                   const uint32_t *offsets, uint32_t n_rows, uint32_t F, const uint32_t *n_rows_dev, uint64_t in_pitch,
                   uint64_t out_pitch, const float *e1, const float *e2, const float *va, const float *vb, bool packed_index, int shape) {
  SegEpilogue epi{e1, e2, va, vb};
  if (n_rows == 0) return NB_OK;
  if (!in_pitch) in_pitch = F;
  if (!out_pitch) out_pitch = F;
  uint32_t fe = F;
  int vec = (push || e1) ? nb_pick_vec(F, in, in_pitch, out, out_pitch) : nb_pick_vec(F, in, in_pitch, out, out_pitch, &fe);
  if (e1 && vec > 1 && (((uintptr_t)va | (uintptr_t)vb) % (4 * vec))) vec = 1;  // epilogue vectors must allow the same vector loads
  if (vec == 4) return launch_segment<4>(ctx, push, in, out, w, idx, offsets, n_rows, fe, n_rows_dev, in_pitch, out_pitch, epi, packed_index, shape);
  if (vec == 2) return launch_segment<2>(ctx, push, in, out, w, idx, offsets, n_rows, fe, n_rows_dev, in_pitch, out_pitch, epi, packed_index, shape);
  return launch_segment<1>(ctx, push, in, out, w, idx, offsets, n_rows, fe, n_rows_dev, in_pitch, out_pitch, epi, packed_index, shape);
}

static int run_segment(nb_ctx *ctx, bool push, const float *in, float *out, const float *w, const uint32_t *idx,
                       const uint32_t *offsets, uint32_t n_rows, uint32_t F, const uint32_t *n_rows_dev = nullptr,
                       uint64_t in_pitch = 0, uint64_t out_pitch = 0, int shape = 0) {
  return nb_run_segment(ctx, push, in, out, w, idx, offsets, n_rows, F, n_rows_dev, in_pitch, out_pitch, nullptr, nullptr, nullptr, nullptr, false, shape);
}

extern "C" {

int nb_aggregate_csc_fwd(nb_ctx *ctx, const float *input, float *output, const float *weight_forward,
                         const uint32_t *row_indices, const uint32_t *column_offset, uint32_t n_dst, uint32_t n_src,
                         uint32_t feature_size) {
  NB_REQUIRE(ctx && (n_dst == 0 || ((input || n_src == 0) && output && column_offset)), NB_ERR_ARG, "nb_aggregate_csc_fwd: NULL argument");
  NB_REQUIRE(feature_size > 0, NB_ERR_ARG, "feature_size must be > 0");
  (void)n_src;
  NB_GUARD(ctx);
  return run_segment(ctx, false, input, output, weight_forward, row_indices, column_offset, n_dst, feature_size);
}

int nb_aggregate_csr_bwd(nb_ctx *ctx, const float *input, float *output, const float *weight_backward,
                         const uint32_t *row_offset, const uint32_t *column_indices, uint32_t n_src, uint32_t n_dst,
                         uint32_t feature_size) {
  NB_REQUIRE(ctx && (n_src == 0 || ((input || n_dst == 0) && output && row_offset)), NB_ERR_ARG, "nb_aggregate_csr_bwd: NULL argument");
  NB_REQUIRE(feature_size > 0, NB_ERR_ARG, "feature_size must be > 0");
  (void)n_dst;
  NB_GUARD(ctx);
  return run_segment(ctx, false, input, output, weight_backward, column_indices, row_offset, n_src, feature_size, nullptr, 0, 0, NB_SEG_SHORT_ROWS);
}

// Extents from device memory (nb_sampler_sizes_dev): no host round trip between sampling and aggregation.
int nb_aggregate_csc_fwd_dyn(nb_ctx *ctx, const float *input, float *output, const float *weight_forward,
                             const uint32_t *row_indices, const uint32_t *column_offset, const uint32_t *n_dst_dev,
                             uint32_t max_dst, uint32_t feature_size, uint32_t input_pitch, uint32_t output_pitch) {
  NB_REQUIRE(ctx && (max_dst == 0 || (output && column_offset)), NB_ERR_ARG, "nb_aggregate_csc_fwd_dyn: NULL argument");
  NB_REQUIRE(feature_size > 0 && input_pitch >= feature_size && output_pitch >= feature_size, NB_ERR_ARG, "bad feature_size / pitch");
  NB_GUARD(ctx);
  return run_segment(ctx, false, input, output, weight_forward, row_indices, column_offset, max_dst, feature_size, n_dst_dev, input_pitch, output_pitch);
}

int nb_aggregate_csr_bwd_dyn(nb_ctx *ctx, const float *input, float *output, const float *weight_backward,
                             const uint32_t *row_offset, const uint32_t *column_indices, const uint32_t *n_src_dev,
                             uint32_t max_src, uint32_t feature_size, uint32_t input_pitch, uint32_t output_pitch) {
  NB_REQUIRE(ctx && (max_src == 0 || (output && row_offset)), NB_ERR_ARG, "nb_aggregate_csr_bwd_dyn: NULL argument");
  NB_REQUIRE(feature_size > 0 && input_pitch >= feature_size && output_pitch >= feature_size, NB_ERR_ARG, "bad feature_size / pitch");
  NB_GUARD(ctx);
  return run_segment(ctx, false, input, output, weight_backward, column_indices, row_offset, max_src, feature_size, n_src_dev, input_pitch, output_pitch,
                     NB_SEG_SHORT_ROWS);
}

// The bottom hop fused with the feature gather (FastSampler::load_feature_gpu + SingleGPU[All]SampleGraphOp::forward in one kernel)
int nb_aggregate_gathered_fwd_dyn(nb_ctx *ctx, const float *table, uint32_t table_pitch, const uint32_t *gather_index, float *output,
                                  const float *weight_forward, const uint32_t *column_offset, const uint32_t *n_dst_dev,
                                  uint32_t max_dst, uint32_t feature_size, uint32_t output_pitch) {
  NB_REQUIRE(ctx && (max_dst == 0 || (table && gather_index && output && column_offset)), NB_ERR_ARG, "nb_aggregate_gathered_fwd_dyn: NULL argument");
  NB_REQUIRE(feature_size > 0 && table_pitch >= feature_size && output_pitch >= feature_size, NB_ERR_ARG, "bad feature_size / pitch");
  NB_GUARD(ctx);
  table = (const float *)nb_mirror_host(ctx, table);
  return nb_run_segment(ctx, false, table, output, weight_forward, gather_index, column_offset, max_dst, feature_size, n_dst_dev, table_pitch,
                        output_pitch, nullptr, nullptr, nullptr, nullptr, true);
}

int nb_aggregate_push_bwd(nb_ctx *ctx, const float *input, float *output, const float *weight, const uint32_t *row_indices,
                          const uint32_t *column_offset, uint32_t n_dst, uint32_t n_src, uint32_t feature_size) {
  NB_REQUIRE(ctx && (n_dst == 0 || (input && row_indices && column_offset)) && (n_src == 0 || output), NB_ERR_ARG,
             "nb_aggregate_push_bwd: NULL argument");
  NB_REQUIRE(feature_size > 0, NB_ERR_ARG, "feature_size must be > 0");
  NB_GUARD(ctx);
  if (n_src) NB_CUDA(cudaMemsetAsync(output, 0, (size_t)n_src * feature_size * sizeof(float), ctx->stream));
  return run_segment(ctx, true, input, output, weight, row_indices, column_offset, n_dst, feature_size);
}

}  // extern "C"
