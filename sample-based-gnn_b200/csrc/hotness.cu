// hotness.cu -- hotness-aware pre-sampling on the GPU.
//
// Replaces nts::op::get_most_neighbor / preSample (core/ntsBaseOp.hpp:333-399, 415-470; per-toolkit copy
// toolkits/GS_SAMPLE_CACHE.hpp:777-849), which the reference runs on the CPU once per super-batch: O(E) atomic count
// propagation over the FULL graph plus a std::sort of |V| counts. Here:
//   k_hot_propagate : counts pushed one hop along the in-edges (integer atomics -> exact, deterministic)
//   k_hot_hist x4 + k_hot_pick : 8-bit MSB radix select of the pivot = count at descending rank cache_num
//                                (no sort), nnz counted in the first pass
//   k_scan<HotOp>   : ids with count >= pivot, ascending, first cache_num of them (the serial order of the reference loop)
// HBM bound: |V|*4 bytes per pass; frontier-proportional adjacency reads.
#include "scan.cuh"

struct HotState {
  uint32_t prefix, k_rem, total, cache_num, pivot, nnz, pad0, pad1;
  uint32_t hist[256];
};

__global__ void k_hot_mark(uint32_t *cnt, const uint32_t *__restrict__ seeds, uint32_t n) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) cnt[seeds[i]] = 1u;
}

// lane l of a warp inspects vertex 32*w + l; vertices with a non-zero count are then expanded by the whole warp
__global__ void __launch_bounds__(256)
k_hot_propagate(const uint32_t *__restrict__ oldc, uint32_t *__restrict__ newc, const uint32_t *__restrict__ col_off,
                const uint32_t *__restrict__ row_idx, uint32_t V) {
  const unsigned lane = lane_id(), warps = (gridDim.x * blockDim.x) >> 5;
  const unsigned groups = (V + 31) / 32;
  for (unsigned w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < groups; w += warps) {
    const unsigned v = w * 32 + lane;
    const uint32_t c = v < V ? oldc[v] : 0u;
    unsigned live = __ballot_sync(FULL_MASK, c > 0);
    while (live) {
      const int src_lane = __ffs(live) - 1;
      live &= live - 1;
      const uint32_t cc = __shfl_sync(FULL_MASK, c, src_lane);
      const unsigned vv = w * 32 + src_lane;
      const uint32_t b = col_off[vv], e = col_off[vv + 1];
      for (uint32_t j = b + lane; j < e; j += 32) atomicAdd(&newc[row_idx[j]], cc);
    }
  }
}

// histogram of byte `shift/8` over the counts whose higher bytes equal st->prefix; pass 0 also counts the non-zeros
__global__ void __launch_bounds__(256)
k_hot_hist(const uint32_t *__restrict__ cnt, uint32_t V, HotState *st, int shift) {
  __shared__ uint32_t h[256];
  __shared__ uint32_t s_nnz;
  h[threadIdx.x] = 0;
  if (threadIdx.x == 0) s_nnz = 0;
  __syncthreads();
  const uint32_t prefix = st->prefix;
  const uint32_t hi_mask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
  uint32_t nnz = 0;
  for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
    const uint32_t c = cnt[v];
    nnz += c > 0;
    if ((c & hi_mask) == (prefix & hi_mask)) atomicAdd(&h[(c >> shift) & 255u], 1u);
  }
  if (shift == 24 && nnz) atomicAdd(&s_nnz, nnz);
  __syncthreads();
  if (h[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], h[threadIdx.x]);
  if (shift == 24 && threadIdx.x == 0 && s_nnz) atomicAdd(&st->nnz, s_nnz);
}

// one block: pick the bin that holds descending rank k_rem, extend the prefix, clear the histogram for the next pass
__global__ void k_hot_pick(HotState *st, uint32_t V, float cache_rate, int shift, uint32_t fixed_cache_num) {
  if (threadIdx.x != 0) return;
  if (shift == 24) {
    // total_sample_num = index of the first zero in the descending order + 1 (core/ntsBaseOp.hpp:366-371)
    uint32_t total = st->nnz < V ? st->nnz + 1 : V;
    uint32_t cache_num = fixed_cache_num != 0xffffffffu ? fixed_cache_num : (uint32_t)((float)total * cache_rate);
    if (cache_num >= V) cache_num = V - 1;
    st->total = total; st->cache_num = cache_num; st->k_rem = cache_num; st->prefix = 0;
  }
  uint32_t k = st->k_rem, acc = 0;
  int b = 255;
  for (; b > 0; b--) {
    if (acc + st->hist[b] > k) break;
    acc += st->hist[b];
  }
  st->k_rem = k - acc;
  st->prefix |= (uint32_t)b << shift;
  if (shift == 0) st->pivot = st->prefix;
  for (int i = 0; i < 256; i++) st->hist[i] = 0;
}

struct HotOp {
  const uint32_t *cnt;
  uint32_t *ids;
  const HotState *st;
  uint32_t V, cap;
  __device__ unsigned n() const { return V; }
  __device__ unsigned load(unsigned v) const { return cnt[v] >= st->pivot ? 1u : 0u; }
  __device__ void store(unsigned v, unsigned excl, unsigned flag) const {
    if (flag && excl < st->cache_num && excl < cap) ids[excl] = v;
  }
  __device__ void total(unsigned) const {}
};

__global__ void k_set_cache_index(uint32_t *cache_map, uint32_t *cache_location, uint32_t super_batch_id, const uint32_t *__restrict__ ids, uint32_t n) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    cache_map[ids[i]] = super_batch_id;
    cache_location[ids[i]] = i;
  }
}

extern "C" {

int nb_set_cache_index(nb_ctx *ctx, uint32_t *cache_map_dev, uint32_t *cache_location_dev, uint32_t super_batch_id,
                       const uint32_t *cache_ids_dev, uint32_t n) {
  NB_REQUIRE(ctx && (n == 0 || (cache_map_dev && cache_location_dev && cache_ids_dev)), NB_ERR_ARG, "nb_set_cache_index: NULL argument");
  NB_GUARD(ctx);
  if (!n) return NB_OK;
  k_set_cache_index<<<nb_grid(n, 256, 4), 256, 0, ctx->stream>>>(cache_map_dev, cache_location_dev, super_batch_id, cache_ids_dev, n);
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
}

int nb_hotness(nb_ctx *ctx, nb_graph *g, const uint32_t *seeds, uint32_t n_seeds, int seeds_on_device, int layers, float cache_rate,
               uint32_t fixed_cache_num, uint32_t *cache_ids_dev, uint32_t capacity, uint32_t *cache_num_out, uint32_t *counts_dev_or_null) {
  NB_REQUIRE(ctx && g && (seeds || n_seeds == 0) && cache_ids_dev && cache_num_out, NB_ERR_ARG, "nb_hotness: NULL argument");
  NB_REQUIRE(layers >= 1, NB_ERR_ARG, "nb_hotness: layers must be >= 1");
  NB_GUARD(ctx);
  const uint32_t V = g->V;
  const size_t n_tiles = (V + SCAN_TILE - 1) / SCAN_TILE + 1;
  uint8_t *p;
  const size_t bytes = 4096 + n_tiles * 8 + ((size_t)V + 64) * 8 + (size_t)n_seeds * 4 + 256;
  int rc = nb_ctx_scratch(ctx, bytes, (void **)&p);
  if (rc) return rc;
  HotState *st = (HotState *)p;
  BatchParams *params = (BatchParams *)(p + 2048);
  unsigned long long *tiles = (unsigned long long *)(p + 4096);
  uint32_t *a = (uint32_t *)(p + 4096 + n_tiles * 8), *b = a + V + 32, *seeds_dev = b + V + 32;
  cudaStream_t s = ctx->stream;
  NB_CUDA(cudaMemsetAsync(p, 0, 4096 + n_tiles * 8 + ((size_t)V + 64) * 8, s));
  BatchParams hp;
  memset(&hp, 0, sizeof(hp));
  hp.epoch = 1;
  NB_CUDA(cudaMemcpyAsync(params, &hp, sizeof(hp), cudaMemcpyHostToDevice, s));
  if (n_seeds) {
    if (!seeds_on_device) {
      NB_CUDA(cudaMemcpyAsync(seeds_dev, seeds, (size_t)n_seeds * 4, cudaMemcpyHostToDevice, s));
      seeds = seeds_dev;
    }
    k_hot_mark<<<nb_grid(n_seeds, 256, 4), 256, 0, s>>>(a, seeds, n_seeds);
    NB_LAUNCH_CHECK(ctx);
  }
  uint32_t *oldc = a, *newc = b;
  for (int layer = 1; layer < layers; layer++) {
    if (layer != 1) {
      uint32_t *t = oldc; oldc = newc; newc = t;
      NB_CUDA(cudaMemsetAsync(newc, 0, (size_t)V * 4, s));
    }
    k_hot_propagate<<<nb_grid((V + 31) / 32, 8, 8), 256, 0, s>>>(oldc, newc, g->col_off, g->row_idx, V);
    NB_LAUNCH_CHECK(ctx);
  }
  const uint32_t *cnt = layers > 1 ? newc : oldc;  // with a single layer nothing is pushed: the reference's new_count stays zero
  if (layers == 1) cnt = newc;
  for (int shift = 24; shift >= 0; shift -= 8) {
    k_hot_hist<<<nb_grid(V, 256, 4), 256, 0, s>>>(cnt, V, st, shift);
    NB_LAUNCH_CHECK(ctx);
    k_hot_pick<<<1, 32, 0, s>>>(st, V, cache_rate, shift, fixed_cache_num);
    NB_LAUNCH_CHECK(ctx);
  }
  HotOp op{cnt, cache_ids_dev, st, V, capacity};
  ScanWs ws = nb_scan_ws(tiles, n_tiles, params);
  k_scan<HotOp><<<nb_grid(V, SCAN_TILE, 4), SCAN_THREADS, 0, s>>>(op, ws);
  NB_LAUNCH_CHECK(ctx);
  if (counts_dev_or_null) NB_CUDA(cudaMemcpyAsync(counts_dev_or_null, cnt, (size_t)V * 4, cudaMemcpyDeviceToDevice, s));
  NB_CUDA(cudaMemcpyAsync(cache_num_out, &st->cache_num, 4, cudaMemcpyDeviceToHost, s));
  NB_CUDA(cudaStreamSynchronize(s));
  NB_REQUIRE(*cache_num_out <= capacity, NB_ERR_CAPACITY, "nb_hotness: %u hot ids exceed the output capacity %u", *cache_num_out, capacity);
  return NB_OK;
}

}  // extern "C"
