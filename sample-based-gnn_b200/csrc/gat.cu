// gat.cu -- GAT edge operators over the sampled CSC.
//
// Legacy-shaped kernels (one per reference kernel, cuda/ntsCUDADistKernel.cuh):
//   scatter_src_dst_to_msg_map :174-196   gather_msg_to_src_dst_map :81-99
//   get_node_max + edge_softmax_forward_norm_block :371-388, :318-368
//   edge_softmax_backward_block :440-484  gather_msg_to_dst :218-232  scatter_dst_to_msg :119-133
// Fused layer: replaces the 5-kernel chain + [E,2F] and [E,F] intermediates of
// toolkits/GAT_SAMPLE_ALL_MULTI.hpp:383-464 (BatchGPUSrcDstScatterOp -> Linear(2F->1) -> leaky_relu ->
// BatchGPUEdgeSoftMax -> mul -> BatchGPUAggregateDst, core/ntsPushdownGraphOp.hpp:490-747):
//   forward : k_gat_node_scores (S*4F read)  +  k_gat_fwd (E*(4 + 4F) + V*4F + 8E, softmax in-warp)
//   backward: k_gat_bwd_edge (E*4F + V*4F)   +  k_gat_bwd_src (CSR gather, E*4F + S*4F, no atomics on dh)
//             + k_gat_bwd_att (S*4F + V*4F, block partial sums -> 2F atomics)
// HBM bound; exp via expf (same as the reference's exp() on float).
#include "common.cuh"

constexpr int GAT_THREADS = 256;

// ---- legacy-shaped ops ---------------------------------------------------------------------
__global__ void __launch_bounds__(GAT_THREADS)
k_scatter_src_dst(float *__restrict__ msg, const float *__restrict__ x, const uint32_t *__restrict__ row_indices,
                  const uint32_t *__restrict__ col_off, const uint32_t *__restrict__ dl, uint32_t n_dst, uint32_t F) {
  const unsigned lane = lane_id(), warp = (blockIdx.x * GAT_THREADS + threadIdx.x) >> 5, warps = (gridDim.x * GAT_THREADS) >> 5;
  for (unsigned d = warp; d < n_dst; d += warps) {
    const float *xd = x + (uint64_t)dl[d] * F;
    for (uint32_t e = col_off[d]; e < col_off[d + 1]; e++) {
      const float *xs = x + (uint64_t)row_indices[e] * F;
      float *m = msg + (uint64_t)e * 2 * F;
      for (unsigned k = lane; k < F; k += 32) { m[k] = xs[k]; m[F + k] = xd[k]; }
    }
  }
}

__global__ void __launch_bounds__(GAT_THREADS)
k_gather_src_dst(float *__restrict__ dx, const float *__restrict__ dmsg, const uint32_t *__restrict__ row_indices,
                 const uint32_t *__restrict__ col_off, const uint32_t *__restrict__ dl, uint32_t n_dst, uint32_t F) {
  const unsigned lane = lane_id(), warp = (blockIdx.x * GAT_THREADS + threadIdx.x) >> 5, warps = (gridDim.x * GAT_THREADS) >> 5;
  for (unsigned d = warp; d < n_dst; d += warps) {
    float *gd = dx + (uint64_t)dl[d] * F;
    const uint32_t beg = col_off[d], end = col_off[d + 1];
    for (unsigned k = lane; k < F; k += 32) {
      float acc = 0.f;
      for (uint32_t e = beg; e < end; e++) {
        const float *m = dmsg + (uint64_t)e * 2 * F;
        atomicAdd(dx + (uint64_t)row_indices[e] * F + k, m[k]);
        acc += m[F + k];
      }
      if (end > beg) atomicAdd(gd + k, acc);
    }
  }
}

__global__ void __launch_bounds__(GAT_THREADS)
k_edge_softmax_fwd(float *__restrict__ out, const float *__restrict__ in, float *__restrict__ cached,
                   const uint32_t *__restrict__ col_off, uint32_t n_dst) {
  const unsigned lane = lane_id(), warp = (blockIdx.x * GAT_THREADS + threadIdx.x) >> 5, warps = (gridDim.x * GAT_THREADS) >> 5;
  for (unsigned d = warp; d < n_dst; d += warps) {
    const uint32_t beg = col_off[d], end = col_off[d + 1];
    if (beg == end) continue;
    float mx = -INFINITY;
    for (uint32_t e = beg + lane; e < end; e += 32) mx = fmaxf(mx, in[e]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (uint32_t e = beg + lane; e < end; e += 32) sum += expf(in[e] - mx);
    sum = warp_sum(sum);
    for (uint32_t e = beg + lane; e < end; e += 32) {
      float a = expf(in[e] - mx) / sum;
      out[e] = a;
      if (cached) cached[e] = a;
    }
  }
}

__global__ void __launch_bounds__(GAT_THREADS)
k_edge_softmax_bwd(float *__restrict__ din, const float *__restrict__ dout, const float *__restrict__ cached,
                   const uint32_t *__restrict__ col_off, uint32_t n_dst) {
  const unsigned lane = lane_id(), warp = (blockIdx.x * GAT_THREADS + threadIdx.x) >> 5, warps = (gridDim.x * GAT_THREADS) >> 5;
  for (unsigned d = warp; d < n_dst; d += warps) {
    const uint32_t beg = col_off[d], end = col_off[d + 1];
    float agg = 0.f;
    for (uint32_t e = beg + lane; e < end; e += 32) agg += dout[e] * cached[e];
    agg = warp_sum(agg);
    for (uint32_t e = beg + lane; e < end; e += 32) din[e] = dout[e] * cached[e] - agg * cached[e];
  }
}

__global__ void __launch_bounds__(GAT_THREADS)
k_gather_msg_to_dst(float *__restrict__ y, const float *__restrict__ msg, const uint32_t *__restrict__ col_off, uint32_t n_dst, uint32_t F) {
  const unsigned lane = lane_id(), warp = (blockIdx.x * GAT_THREADS + threadIdx.x) >> 5, warps = (gridDim.x * GAT_THREADS) >> 5;
  for (unsigned d = warp; d < n_dst; d += warps) {
    const uint32_t beg = col_off[d], end = col_off[d + 1];
    for (unsigned k = lane; k < F; k += 32) {
      float acc = 0.f;
      for (uint32_t e = beg; e < end; e++) acc += msg[(uint64_t)e * F + k];
      y[(uint64_t)d * F + k] = acc;
    }
  }
}

__global__ void __launch_bounds__(GAT_THREADS)
k_scatter_dst_to_msg(float *__restrict__ msg, const float *__restrict__ y, const uint32_t *__restrict__ col_off, uint32_t n_dst, uint32_t F) {
  const unsigned lane = lane_id(), warp = (blockIdx.x * GAT_THREADS + threadIdx.x) >> 5, warps = (gridDim.x * GAT_THREADS) >> 5;
  for (unsigned d = warp; d < n_dst; d += warps) {
    const uint32_t beg = col_off[d], end = col_off[d + 1];
    for (unsigned k = lane; k < F; k += 32) {
      float v = y[(uint64_t)d * F + k];
      for (uint32_t e = beg; e < end; e++) msg[(uint64_t)e * F + k] = v;
    }
  }
}

// ---- fused layer ---------------------------------------------------------------------------
// al[s] = h[s,:].att[0:F], ar[s] = h[s,:].att[F:2F]. One warp per row; rows of 4k floats move as 128-bit vectors with ROWS rows in
// flight per warp (one 512-byte row is a single vector per lane: latency bound unless several rows overlap), att stays in registers.
template <int VEC, int ROWS>
__global__ void __launch_bounds__(GAT_THREADS)
k_gat_node_scores(const float *__restrict__ h, const float *__restrict__ att, float *__restrict__ al, float *__restrict__ ar,
                  uint32_t n_src, uint32_t F) {
  const unsigned lane = lane_id(), warp = (blockIdx.x * GAT_THREADS + threadIdx.x) >> 5, warps = (gridDim.x * GAT_THREADS) >> 5;
  const unsigned nvec = F / VEC;
  if (nvec <= 32) {   // the whole row is one vector per lane
    Vec<VEC> wa, wb;
    wa.zero(); wb.zero();
    if (lane < nvec) { wa.load_cached(att + (uint64_t)lane * VEC); wb.load_cached(att + F + (uint64_t)lane * VEC); }
    for (unsigned s0 = warp * ROWS; s0 < n_src; s0 += warps * ROWS) {
      Vec<VEC> x[ROWS];
#pragma unroll
      for (int r = 0; r < ROWS; r++) {
        x[r].zero();
        if (s0 + r < n_src && lane < nvec) x[r].load(h + (uint64_t)(s0 + r) * F + (uint64_t)lane * VEC);
      }
#pragma unroll
      for (int r = 0; r < ROWS; r++) {
        const float a = warp_sum(x[r].dot(wa)), b = warp_sum(x[r].dot(wb));
        if (lane == 0 && s0 + r < n_src) { al[s0 + r] = a; ar[s0 + r] = b; }
      }
    }
    return;
  }
  for (unsigned s = warp; s < n_src; s += warps) {
    const float *p = h + (uint64_t)s * F;
    float a = 0.f, b = 0.f;
    for (unsigned k = lane; k < nvec; k += 32) {
      Vec<VEC> x, wa, wb;
      x.load(p + (uint64_t)k * VEC);
      wa.load_cached(att + (uint64_t)k * VEC);
      wb.load_cached(att + F + (uint64_t)k * VEC);
      a += x.dot(wa); b += x.dot(wb);
    }
    a = warp_sum(a); b = warp_sum(b);
    if (lane == 0) { al[s] = a; ar[s] = b; }
  }
}

__device__ __forceinline__ float lrelu(float s, float slope) { return s > 0.f ? s : slope * s; }

// One warp per dst column. Columns of <= 32 edges (every sampled column: fanout <= 32) keep the edge's source id and score in
// the lane that owns the edge: ONE gather of al[src] per edge, max / sum by shuffles, weights straight from registers
// (the general path below re-reads al[row_indices[e]] in three passes). UNR input rows are in flight before the first
// accumulate, as in k_segment_reduce; accumulation order is the column's stored order either way.
template <int VEC, int CHUNK, int UNR>
__global__ void __launch_bounds__(GAT_THREADS)
k_gat_fwd(const float *__restrict__ h, const float *__restrict__ al, const float *__restrict__ ar, float slope,
          const uint32_t *__restrict__ col_off, const uint32_t *__restrict__ row_indices, const uint32_t *__restrict__ dl,
          uint32_t n_dst, uint32_t nvec, uint64_t pitch, float *__restrict__ score_pre, float *__restrict__ alpha, float *__restrict__ out) {
  const unsigned lane = lane_id(), warp = (blockIdx.x * GAT_THREADS + threadIdx.x) >> 5, warps = (gridDim.x * GAT_THREADS) >> 5;
  for (unsigned d = warp; d < n_dst; d += warps) {
    const uint32_t beg = col_off[d], end = col_off[d + 1];
    const float ard = (beg < end) ? ar[dl[d]] : 0.f;
    const bool short_col = end - beg <= 32u;
    uint32_t my_idx = 0;
    float my_s = 0.f, mx = -INFINITY, sum = 0.f;
    if (short_col) {
      const bool mine = beg + lane < end;
      if (mine) { my_idx = row_indices[beg + lane]; my_s = al[my_idx] + ard; }
      mx = warp_max(mine ? lrelu(my_s, slope) : -INFINITY);
      sum = warp_sum(mine ? expf(lrelu(my_s, slope) - mx) : 0.f);
    } else {
      for (uint32_t e = beg + lane; e < end; e += 32) mx = fmaxf(mx, lrelu(al[row_indices[e]] + ard, slope));
      mx = warp_max(mx);
      for (uint32_t e = beg + lane; e < end; e += 32) sum += expf(lrelu(al[row_indices[e]] + ard, slope) - mx);
      sum = warp_sum(sum);
    }
    for (unsigned c0 = 0; c0 < nvec; c0 += 32 * CHUNK) {
      Vec<VEC> acc[CHUNK];
#pragma unroll
      for (int c = 0; c < CHUNK; c++) acc[c].zero();
      for (uint32_t j0 = beg; j0 < end; j0 += 32) {
        const uint32_t cnt = min(32u, end - j0);
        float my_w = 0.f;
        if (lane < cnt) {
          if (!short_col) { my_idx = row_indices[j0 + lane]; my_s = al[my_idx] + ard; }
          my_w = expf(lrelu(my_s, slope) - mx) / sum;
          if (c0 == 0) { score_pre[j0 + lane] = my_s; alpha[j0 + lane] = my_w; }
        }
        for (uint32_t t = 0; t < cnt; t += UNR) {
          Vec<VEC> x[UNR][CHUNK];
          float w[UNR];
#pragma unroll
          for (int u = 0; u < UNR; u++) {
            const uint32_t tt = t + u < cnt ? t + u : t;
            const uint32_t s0 = __shfl_sync(FULL_MASK, my_idx, tt);
            w[u] = __shfl_sync(FULL_MASK, my_w, tt);
            const float *p0 = h + (uint64_t)s0 * pitch;
            if (t + u < cnt) {
#pragma unroll
              for (int c = 0; c < CHUNK; c++) {
                const unsigned k = c0 + c * 32 + lane;
                if (k < nvec) x[u][c].load(p0 + (uint64_t)k * VEC);
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
      float *o = out + (uint64_t)d * pitch;
#pragma unroll
      for (int c = 0; c < CHUNK; c++) {
        const unsigned k = c0 + c * 32 + lane;
        if (k < nvec) acc[c].store(o + (uint64_t)k * VEC);
      }
    }
  }
}

// per dst column: da_e = dout[d].h[src_e]; dm_e = alpha_e (da_e - sum_e' alpha_e' da_e'); ds_e = lrelu'(pre_e) dm_e
// writes ds[E] and dsum[V] = sum_e ds_e. The dout row stays in registers; two h rows are in flight per step.
template <int VEC, int CHUNK>
__global__ void __launch_bounds__(GAT_THREADS)
k_gat_bwd_edge(const float *__restrict__ h, const float *__restrict__ dout, const float *__restrict__ score_pre,
               const float *__restrict__ alpha, float slope, const uint32_t *__restrict__ col_off,
               const uint32_t *__restrict__ row_indices, uint32_t n_dst, uint32_t nvec, uint64_t pitch, float *__restrict__ ds,
               float *__restrict__ dsum) {
  const unsigned lane = lane_id(), warp = (blockIdx.x * GAT_THREADS + threadIdx.x) >> 5, warps = (gridDim.x * GAT_THREADS) >> 5;
  for (unsigned d = warp; d < n_dst; d += warps) {
    const uint32_t beg = col_off[d], end = col_off[d + 1];
    if (beg == end) { if (lane == 0) dsum[d] = 0.f; continue; }
    Vec<VEC> g[CHUNK];
#pragma unroll
    for (int c = 0; c < CHUNK; c++) {
      const unsigned k = c * 32 + lane;
      if (k < nvec) g[c].load(dout + (uint64_t)d * pitch + (uint64_t)k * VEC); else g[c].zero();
    }
    float agg = 0.f;
    for (uint32_t j0 = beg; j0 < end; j0 += 32) {
      const uint32_t cnt = min(32u, end - j0);
      uint32_t my_idx = 0;
      float my_alpha = 0.f, my_da = 0.f;
      if (lane < cnt) { my_idx = row_indices[j0 + lane]; my_alpha = alpha[j0 + lane]; }
      for (uint32_t t = 0; t < cnt; t += 2) {
        const uint32_t t1 = t + 1 < cnt ? t + 1 : t;
        const float *p0 = h + (uint64_t)__shfl_sync(FULL_MASK, my_idx, t) * pitch;
        const float *p1 = h + (uint64_t)__shfl_sync(FULL_MASK, my_idx, t1) * pitch;
        float d0 = 0.f, d1 = 0.f;
#pragma unroll
        for (int c = 0; c < CHUNK; c++) {
          const unsigned k = c * 32 + lane;
          if (k < nvec) {
            Vec<VEC> x0, x1;
            x0.load(p0 + (uint64_t)k * VEC);
            x1.load(p1 + (uint64_t)k * VEC);
            d0 += g[c].dot(x0);
            d1 += g[c].dot(x1);
          }
        }
        d0 = warp_sum(d0);
        d1 = warp_sum(d1);
        if (lane == t) my_da = d0;
        if (lane == t1 && t1 != t) my_da = d1;
      }
      agg += warp_sum(lane < cnt ? my_da * my_alpha : 0.f);
      if (lane < cnt) ds[j0 + lane] = my_da;  // stash da until the column total is known
    }
    float tot = 0.f;
    for (uint32_t e = beg + lane; e < end; e += 32) {
      const float dm = alpha[e] * (ds[e] - agg);
      const float v = score_pre[e] > 0.f ? dm : slope * dm;
      ds[e] = v;
      tot += v;
    }
    tot = warp_sum(tot);
    if (lane == 0) dsum[d] = tot;
  }
}

// CSR-order weights and per-src scalars for the segment reduction: w[j] = alpha[e_j], rs[s] = sum_j ds[e_j], dd[s] = dsum of the dst equal to s
__global__ void __launch_bounds__(GAT_THREADS)
k_gat_bwd_rows(const float *__restrict__ alpha, const float *__restrict__ ds, const float *__restrict__ dsum,
               const uint32_t *__restrict__ row_offset, const uint32_t *__restrict__ csr_to_csc, const uint32_t *__restrict__ src_to_dst,
               uint32_t n_src, float *__restrict__ wcsr, float *__restrict__ rs, float *__restrict__ dd) {
  // thread per src: rows of up to four entries (the common case: ~1.6 entries per row) with all their loads in flight at once,
  // rows of up to 32 in a plain loop; a hub row is walked by the whole warp afterwards instead of serialising one thread on
  // hundreds of dependent loads
  const unsigned lane = lane_id();
  const unsigned n_round = (n_src + 31u) & ~31u;
  for (unsigned s = blockIdx.x * blockDim.x + threadIdx.x; s < n_round; s += gridDim.x * blockDim.x) {
    uint32_t beg = 0, end = 0;
    if (s < n_src) { beg = row_offset[s]; end = row_offset[s + 1]; }
    const bool is_long = end - beg > 32u;
    if (s < n_src && !is_long && end - beg > 4u) {
      float r = 0.f;
      for (uint32_t j = beg; j < end; j++) {
        const uint32_t e = csr_to_csc[j];
        wcsr[j] = alpha[e];
        r += ds[e];
      }
      rs[s] = r;
    } else if (s < n_src && !is_long) {
      // up to four entries: all the index loads, then all the value loads, in flight together (a dependent per-entry loop
      // costs two memory round trips per entry)
      uint32_t e[4];
      float a[4], d[4];
#pragma unroll
      for (int k = 0; k < 4; k++) e[k] = beg + k < end ? csr_to_csc[beg + k] : 0u;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        a[k] = beg + k < end ? alpha[e[k]] : 0.f;
        d[k] = beg + k < end ? ds[e[k]] : 0.f;
      }
      float r = 0.f;
#pragma unroll
      for (int k = 0; k < 4; k++)
        if (beg + k < end) { wcsr[beg + k] = a[k]; r += d[k]; }
      rs[s] = r;
    }
    if (s < n_src) {
      const uint32_t d = src_to_dst[s];
      dd[s] = d != 0xffffffffu ? dsum[d] : 0.f;
    }
    unsigned pending = __ballot_sync(FULL_MASK, is_long);
    while (pending) {
      const int owner = __ffs(pending) - 1;
      pending &= pending - 1;
      const uint32_t b = __shfl_sync(FULL_MASK, beg, owner), e_end = __shfl_sync(FULL_MASK, end, owner);
      float r = 0.f;
      for (uint32_t j = b + lane; j < e_end; j += 32) {
        const uint32_t e = csr_to_csc[j];
        wcsr[j] = alpha[e];
        r += ds[e];
      }
      r = warp_sum(r);
      if ((int)lane == owner) rs[s] = r;
    }
  }
}

// datt[0:F] += sum_s rs[s] h[s,:];  datt[F:2F] += sum_d dsum[d] h[dl[d],:]
__global__ void __launch_bounds__(GAT_THREADS)
k_gat_bwd_att(const float *__restrict__ h, const float *__restrict__ rs, const float *__restrict__ dd, uint32_t n_src, uint32_t F,
              float *__restrict__ partial) {
  // one pass over H. A block owns a contiguous slice of rows; thread (g, k) sums feature column k over the rows g, g+G, ... of the
  // slice (G = blockDim / F row groups when F <= blockDim), four rows in flight; the groups are combined in shared memory and
  // the block writes its 2F partial sums for stage 2. dd[s] is the column total of the dst that equals src s (0 if s is not a dst).
  __shared__ float s_a[GAT_THREADS], s_b[GAT_THREADS];
  const unsigned rows_per_block = (n_src + gridDim.x - 1) / gridDim.x;
  const unsigned r0 = min(n_src, blockIdx.x * rows_per_block), r1 = min(n_src, r0 + rows_per_block);   // an empty slice writes zeros
  const unsigned G = F <= GAT_THREADS ? GAT_THREADS / F : 1;
  const unsigned g = threadIdx.x / (F <= GAT_THREADS ? F : GAT_THREADS), lanes = F <= GAT_THREADS ? F : GAT_THREADS;
  for (unsigned k0 = 0; k0 < F; k0 += lanes) {
    const unsigned k = k0 + threadIdx.x % lanes;
    float a = 0.f, b = 0.f;
    if (g < G && k < F) {
      unsigned s = r0 + g;
      for (; s + 3 * G < r1; s += 4 * G) {
        const float x0 = h[(uint64_t)s * F + k], x1 = h[(uint64_t)(s + G) * F + k], x2 = h[(uint64_t)(s + 2 * G) * F + k],
                    x3 = h[(uint64_t)(s + 3 * G) * F + k];
        a += rs[s] * x0 + rs[s + G] * x1 + rs[s + 2 * G] * x2 + rs[s + 3 * G] * x3;
        b += dd[s] * x0 + dd[s + G] * x1 + dd[s + 2 * G] * x2 + dd[s + 3 * G] * x3;
      }
      for (; s < r1; s += G) {
        const float x = h[(uint64_t)s * F + k];
        a += rs[s] * x;
        b += dd[s] * x;
      }
    }
    s_a[threadIdx.x] = a;
    s_b[threadIdx.x] = b;
    __syncthreads();
    if (g == 0 && k < F) {
      for (unsigned gg = 1; gg < G; gg++) { a += s_a[gg * lanes + threadIdx.x]; b += s_b[gg * lanes + threadIdx.x]; }
      partial[(uint64_t)blockIdx.x * 2 * F + k] = a;       // no float atomics: stage 2 sums the block partials in block order
      partial[(uint64_t)blockIdx.x * 2 * F + F + k] = b;
    }
    __syncthreads();
  }
}
// stage 2: datt[k] = sum over the block partials. One warp per output column: lane l adds blocks l, l+32, ... in order, the 32 lane
// sums are combined by a fixed shuffle tree -- a fixed summation order, so the result is deterministic run to run.
__global__ void __launch_bounds__(GAT_THREADS)
k_gat_att_reduce(const float *__restrict__ partial, uint32_t n_blocks, uint32_t n2f, float *__restrict__ datt) {
  const unsigned lane = lane_id(), k = (blockIdx.x * GAT_THREADS + threadIdx.x) >> 5;
  if (k >= n2f) return;
  float acc = 0.f;
  for (uint32_t b = lane; b < n_blocks; b += 32) acc += partial[(uint64_t)b * n2f + k];
  acc = warp_sum(acc);
  if (lane == 0) datt[k] = acc;
}

template <int VEC>
static int launch_gat_fwd(nb_ctx *ctx, const float *h, const float *al, const float *ar, float slope, const uint32_t *co,
                          const uint32_t *ri, const uint32_t *dl, uint32_t n_dst, uint32_t F, float *pre, float *alpha, float *out) {
  const uint32_t nvec = F / VEC, per_lane = (nvec + 31) / 32;
  const unsigned grid = nb_grid(n_dst, GAT_THREADS / 32, 8);
#define NB_GAT(C) k_gat_fwd<VEC, C, (C <= 2 ? 4 : 2)><<<grid, GAT_THREADS, 0, ctx->stream>>>(h, al, ar, slope, co, ri, dl, n_dst, nvec, F, pre, alpha, out)
  if (per_lane <= 1) NB_GAT(1); else if (per_lane <= 2) NB_GAT(2); else if (per_lane <= 4) NB_GAT(4);
  else if (per_lane <= 8) NB_GAT(8); else NB_GAT(12);
#undef NB_GAT
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
}

extern "C" {

int nb_scatter_src_dst_to_msg(nb_ctx *ctx, float *message, const float *src_feature, const uint32_t *row_indices,
                              const uint32_t *column_offset, uint32_t n_dst, uint32_t feature_size, const uint32_t *dst_local_id) {
  NB_REQUIRE(ctx && (n_dst == 0 || (message && src_feature && row_indices && column_offset && dst_local_id)), NB_ERR_ARG, "nb_scatter_src_dst_to_msg: NULL argument");
  NB_GUARD(ctx);
  if (n_dst == 0) return NB_OK;
  k_scatter_src_dst<<<nb_grid(n_dst, GAT_THREADS / 32, 8), GAT_THREADS, 0, ctx->stream>>>(message, src_feature, row_indices, column_offset, dst_local_id, n_dst, feature_size);
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
}

int nb_gather_msg_to_src_dst(nb_ctx *ctx, float *src_grad, const float *message_grad, const uint32_t *row_indices,
                             const uint32_t *column_offset, uint32_t n_dst, uint32_t n_src, uint32_t feature_size,
                             const uint32_t *dst_local_id) {
  NB_REQUIRE(ctx && (n_dst == 0 || (src_grad && message_grad && row_indices && column_offset && dst_local_id)), NB_ERR_ARG, "nb_gather_msg_to_src_dst: NULL argument");
  NB_GUARD(ctx);
  if (n_src) NB_CUDA(cudaMemsetAsync(src_grad, 0, (size_t)n_src * feature_size * 4, ctx->stream));
  if (n_dst == 0) return NB_OK;
  k_gather_src_dst<<<nb_grid(n_dst, GAT_THREADS / 32, 8), GAT_THREADS, 0, ctx->stream>>>(src_grad, message_grad, row_indices, column_offset, dst_local_id, n_dst, feature_size);
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
}

int nb_edge_softmax_fwd(nb_ctx *ctx, float *msg_output, const float *msg_input, float *msg_cached, const uint32_t *column_offset, uint32_t n_dst) {
  NB_REQUIRE(ctx && (n_dst == 0 || (msg_output && msg_input && column_offset)), NB_ERR_ARG, "nb_edge_softmax_fwd: NULL argument");
  NB_GUARD(ctx);
  if (n_dst == 0) return NB_OK;
  k_edge_softmax_fwd<<<nb_grid(n_dst, GAT_THREADS / 32, 8), GAT_THREADS, 0, ctx->stream>>>(msg_output, msg_input, msg_cached, column_offset, n_dst);
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
}

int nb_edge_softmax_bwd(nb_ctx *ctx, float *msg_input_grad, const float *msg_output_grad, const float *msg_cached, const uint32_t *column_offset, uint32_t n_dst) {
  NB_REQUIRE(ctx && (n_dst == 0 || (msg_input_grad && msg_output_grad && msg_cached && column_offset)), NB_ERR_ARG, "nb_edge_softmax_bwd: NULL argument");
  NB_GUARD(ctx);
  if (n_dst == 0) return NB_OK;
  k_edge_softmax_bwd<<<nb_grid(n_dst, GAT_THREADS / 32, 8), GAT_THREADS, 0, ctx->stream>>>(msg_input_grad, msg_output_grad, msg_cached, column_offset, n_dst);
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
}

int nb_gather_msg_to_dst(nb_ctx *ctx, float *dst_feature, const float *message, const uint32_t *column_offset, uint32_t n_dst, uint32_t feature_size) {
  NB_REQUIRE(ctx && (n_dst == 0 || (dst_feature && message && column_offset)), NB_ERR_ARG, "nb_gather_msg_to_dst: NULL argument");
  NB_GUARD(ctx);
  if (n_dst == 0) return NB_OK;
  k_gather_msg_to_dst<<<nb_grid(n_dst, GAT_THREADS / 32, 8), GAT_THREADS, 0, ctx->stream>>>(dst_feature, message, column_offset, n_dst, feature_size);
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
}

int nb_scatter_dst_to_msg(nb_ctx *ctx, float *message, const float *dst_feature, const uint32_t *column_offset, uint32_t n_dst, uint32_t feature_size) {
  NB_REQUIRE(ctx && (n_dst == 0 || (message && dst_feature && column_offset)), NB_ERR_ARG, "nb_scatter_dst_to_msg: NULL argument");
  NB_GUARD(ctx);
  if (n_dst == 0) return NB_OK;
  k_scatter_dst_to_msg<<<nb_grid(n_dst, GAT_THREADS / 32, 8), GAT_THREADS, 0, ctx->stream>>>(message, dst_feature, column_offset, n_dst, feature_size);
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
}

int nb_gat_fwd(nb_ctx *ctx, const float *h, const float *att, float negative_slope, const uint32_t *column_offset,
               const uint32_t *row_indices, const uint32_t *dst_local_id, uint32_t n_dst, uint32_t n_src,
               uint32_t feature_size, float *score_pre, float *alpha, float *out) {
  NB_REQUIRE(ctx && h && att && column_offset && row_indices && dst_local_id && score_pre && alpha && out, NB_ERR_ARG, "nb_gat_fwd: NULL argument");
  NB_REQUIRE(feature_size > 0, NB_ERR_ARG, "feature_size must be > 0");
  NB_GUARD(ctx);
  if (n_dst == 0) return NB_OK;
  float *scratch;
  int rc = nb_ctx_scratch(ctx, (size_t)n_src * 2 * sizeof(float), (void **)&scratch);
  if (rc) return rc;
  float *al = scratch, *ar = scratch + n_src;
  int vec = nb_pick_vec(feature_size, h, feature_size, out, feature_size);
  {
    const int av = (((uintptr_t)att | (uintptr_t)(att + feature_size)) % 16 == 0 && vec == 4) ? 4 : (((uintptr_t)att | (uintptr_t)(att + feature_size)) % 8 == 0 && vec >= 2 ? 2 : 1);
    const unsigned grid = nb_grid(n_src, GAT_THREADS / 32 * 4, 8);
    if (av == 4) k_gat_node_scores<4, 4><<<grid, GAT_THREADS, 0, ctx->stream>>>(h, att, al, ar, n_src, feature_size);
    else if (av == 2) k_gat_node_scores<2, 4><<<grid, GAT_THREADS, 0, ctx->stream>>>(h, att, al, ar, n_src, feature_size);
    else k_gat_node_scores<1, 4><<<grid, GAT_THREADS, 0, ctx->stream>>>(h, att, al, ar, n_src, feature_size);
    NB_LAUNCH_CHECK(ctx);
  }
  if (vec == 4) return launch_gat_fwd<4>(ctx, h, al, ar, negative_slope, column_offset, row_indices, dst_local_id, n_dst, feature_size, score_pre, alpha, out);
  if (vec == 2) return launch_gat_fwd<2>(ctx, h, al, ar, negative_slope, column_offset, row_indices, dst_local_id, n_dst, feature_size, score_pre, alpha, out);
  return launch_gat_fwd<1>(ctx, h, al, ar, negative_slope, column_offset, row_indices, dst_local_id, n_dst, feature_size, score_pre, alpha, out);
}

int nb_gat_bwd(nb_ctx *ctx, const float *h, const float *att, float negative_slope, const float *dout,
               const float *score_pre, const float *alpha, const uint32_t *column_offset, const uint32_t *row_indices,
               const uint32_t *dst_local_id, const uint32_t *row_offset, const uint32_t *column_indices,
               const uint32_t *csr_to_csc, const uint32_t *src_to_dst, uint32_t n_dst, uint32_t n_src, uint32_t n_edges,
               uint32_t feature_size, float *dh, float *datt) {
  NB_REQUIRE(ctx && h && att && dout && score_pre && alpha && column_offset && row_indices && dst_local_id && row_offset &&
                 column_indices && csr_to_csc && src_to_dst && dh && datt, NB_ERR_ARG, "nb_gat_bwd: NULL argument (needs a sampler built with MERGE_SRC_DST|BUILD_CSR)");
  NB_GUARD(ctx);
  const uint32_t F = feature_size;
  NB_CUDA(cudaMemsetAsync(datt, 0, (size_t)2 * F * 4, ctx->stream));
  if (n_src == 0) return NB_OK;
  float *scratch;
  const unsigned att_blocks = (unsigned)min((uint64_t)ctx->sm_count * 8, ((uint64_t)n_src + 15) / 16);   // >= 16 rows of H per block
  int rc = nb_ctx_scratch(ctx, ((size_t)2 * n_edges + n_dst + 2 * (size_t)n_src + 64 + (size_t)att_blocks * 2 * F) * sizeof(float), (void **)&scratch);
  if (rc) return rc;
  float *ds = scratch, *wcsr = ds + n_edges, *dsum = wcsr + n_edges, *rs = dsum + n_dst + 8, *dd = rs + n_src + 8, *partial = dd + n_src + 8;
  if (n_dst) {
    const int vec = nb_pick_vec(F, h, F, dout, F);
    const uint32_t nvec = F / vec, per_lane = (nvec + 31) / 32;
    const unsigned grid = nb_grid(n_dst, GAT_THREADS / 32, 8);
    bool done = false;
#define NB_GBE(V, C)                                                                                                              \
  if (!done && vec == V && per_lane <= C) {                                                                                        \
    k_gat_bwd_edge<V, C><<<grid, GAT_THREADS, 0, ctx->stream>>>(h, dout, score_pre, alpha, negative_slope, column_offset, row_indices, \
                                                               n_dst, nvec, F, ds, dsum);                                         \
    done = true;                                                                                                                   \
  }
    NB_GBE(4, 1) NB_GBE(4, 2) NB_GBE(4, 4) NB_GBE(4, 8) NB_GBE(2, 2) NB_GBE(2, 4) NB_GBE(2, 8) NB_GBE(2, 16) NB_GBE(1, 2) NB_GBE(1, 4) NB_GBE(1, 8) NB_GBE(1, 16) NB_GBE(1, 32)
#undef NB_GBE
    NB_REQUIRE(done, NB_ERR_UNSUPPORTED, "nb_gat_bwd: feature_size %u too wide for the fused backward (max 1024)", F);
    NB_LAUNCH_CHECK(ctx);
  } else {
    NB_CUDA(cudaMemsetAsync(ds, 0, (size_t)n_edges * 4, ctx->stream));
  }
  // dh[s,:] = sum_j alpha[e_j] dout[dst_j,:] + rs[s] att[0:F] + dd[s] att[F:2F], rs[s] = sum_j ds[e_j], dd[s] = dsum of the dst equal to s:
  // the tuned CSR segment reduction derives the row's weights and both scalars on the fly (no staging kernel) and leaves rs / dd
  // behind for the attention-gradient pass
  (void)wcsr;
  rc = nb_run_segment_gat(ctx, dout, dh, column_indices, row_offset, n_src, F, csr_to_csc, alpha, ds, dsum, src_to_dst, att, att + F, rs, dd);
  if (rc) return rc;
  k_gat_bwd_att<<<att_blocks, GAT_THREADS, 0, ctx->stream>>>(h, rs, dd, n_src, F, partial);
  NB_LAUNCH_CHECK(ctx);
  k_gat_att_reduce<<<(2 * F * 32 + GAT_THREADS - 1) / GAT_THREADS, GAT_THREADS, 0, ctx->stream>>>(partial, att_blocks, 2 * F, datt);
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
}

}  // extern "C"
