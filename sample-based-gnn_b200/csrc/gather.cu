// gather.cu -- feature / label gather and hot-row override.
//
// Replaces (reference file:line)
//   zero_copy_feature_move_gpu_kernel            cuda/ntsCUDATransferKernel.cuh:97-115  (4-byte loads over PCIe zero-copy)
//   zero_copy_feature_move_gpu_cache_kernel      :154-167, gather_feature_from_gpu_cache_kernel :169-183
//     + the serial CPU hot/cold split            core/ntsFastSampler.hpp:284-298
//   global_copy_label_move_gpu_kernel            :203-212
//   dev_load_share_embedding[_and_feature]_kernel, dev_load_share_aggregate_kernel :412-529
//
// HBM bound. Algorithmic bytes per row: 4 (id) + 2*4F (read row + write row). One warp moves one
// row with 128-/64-bit read-only streaming loads, ROWS_IN_FLIGHT rows per warp iteration so that
// every lane has >= 4 independent 16-byte requests outstanding; grid = resident warps of 148 SMs.
#include "common.cuh"

struct nb_table {
  nb_ctx *ctx;
  uint32_t n_shards, feature_size, pitch;
  uint64_t n_rows;
  const float **shards_dev;  // device array of n_shards row-base pointers
};

constexpr int GATHER_THREADS = 256;

// MODE 0: plain table. MODE 1: hot/cold (cache_map slot != -1 -> cache table). MODE 2: sharded table (v % n, v / n).
// CHUNK = vectors per lane held in registers: a row of up to 32*CHUNK vectors is fetched with CHUNK independent
// requests per lane before the first store (602 floats = 301 float2 -> CHUNK 10, 2.4 KB in flight per warp);
// the next row's id is fetched one iteration ahead so the id -> row dependency is off the critical path.
template <int VEC, int CHUNK, int MODE>
__global__ void __launch_bounds__(GATHER_THREADS)
k_gather_rows(float *__restrict__ out, const float *__restrict__ table, uint64_t table_pitch, const float *__restrict__ cache,
              uint64_t cache_pitch, const uint32_t *__restrict__ cache_map, const float *const *__restrict__ shards,
              uint32_t n_shards, const uint32_t *__restrict__ ids, uint32_t n_rows, const uint32_t *__restrict__ n_rows_dev,
              uint32_t nvec, uint64_t out_pitch, uint32_t *hit_count) {
  const unsigned lane = lane_id();
  const unsigned warp = (blockIdx.x * GATHER_THREADS + threadIdx.x) >> 5;
  const unsigned warps = (gridDim.x * GATHER_THREADS) >> 5;
  if (n_rows_dev) n_rows = min(n_rows, *n_rows_dev);
  unsigned hits = 0;
  uint32_t v_next = warp < n_rows ? ids[warp] : 0;
  for (unsigned i = warp; i < n_rows; i += warps) {
    const uint32_t v = v_next;
    if (i + warps < n_rows) v_next = ids[i + warps];
    const float *src;
    if (MODE == 0) src = table + (uint64_t)v * table_pitch;
    else if (MODE == 1) {
      const uint32_t slot = cache_map[v];
      if (slot != 0xffffffffu) { src = cache + (uint64_t)slot * cache_pitch; hits++; }
      else src = table + (uint64_t)v * table_pitch;
    } else src = shards[v % n_shards] + (uint64_t)(v / n_shards) * table_pitch;
    float *dst = out + (uint64_t)i * out_pitch;
    for (unsigned c0 = 0; c0 < nvec; c0 += 32 * CHUNK) {
      Vec<VEC> x[CHUNK];
#pragma unroll
      for (int c = 0; c < CHUNK; c++) {
        const unsigned k = c0 + c * 32 + lane;
        if (k < nvec) x[c].load(src + (uint64_t)k * VEC);
      }
#pragma unroll
      for (int c = 0; c < CHUNK; c++) {
        const unsigned k = c0 + c * 32 + lane;
        if (k < nvec) x[c].store(dst + (uint64_t)k * VEC);
      }
    }
  }
  if (MODE == 1 && hit_count && lane == 0 && hits) atomicAdd(hit_count, hits);
}

__global__ void k_gather_labels(int64_t *__restrict__ out, const int64_t *__restrict__ labels, const uint32_t *__restrict__ ids, uint32_t n) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = labels[ids[i]];
}

// rows i whose cache_map[destination[i]] matches are overwritten by share[cache_location[...]]
template <int VEC>
__global__ void __launch_bounds__(GATHER_THREADS)
k_row_override(float *__restrict__ out_a, const float *__restrict__ share_a, uint32_t nvec_a, uint32_t fa,
               float *__restrict__ out_b, const float *__restrict__ share_b, uint32_t nvec_b, uint32_t fb,
               const uint32_t *__restrict__ cache_map, const uint32_t *__restrict__ cache_location,
               const uint32_t *__restrict__ destination, uint32_t n_dst, uint32_t super_batch_id) {
  const unsigned lane = lane_id();
  const unsigned warp = (blockIdx.x * GATHER_THREADS + threadIdx.x) >> 5;
  const unsigned warps = (gridDim.x * GATHER_THREADS) >> 5;
  for (unsigned i = warp; i < n_dst; i += warps) {
    const uint32_t v = destination[i];
    const uint32_t f = cache_map[v];
    const bool hit = super_batch_id == 0xffffffffu ? (f != 0xffffffffu) : (f == super_batch_id);
    if (!hit) continue;
    const uint32_t loc = cache_location[v];
    for (unsigned k = lane; k < nvec_a; k += 32) {
      Vec<VEC> a;
      a.load(share_a + (uint64_t)loc * fa + (uint64_t)k * VEC);
      a.store(out_a + (uint64_t)i * fa + (uint64_t)k * VEC);
    }
    if (out_b)
      for (unsigned k = lane; k < nvec_b; k += 32) {
        Vec<VEC> a;
        a.load(share_b + (uint64_t)loc * fb + (uint64_t)k * VEC);
        a.store(out_b + (uint64_t)i * fb + (uint64_t)k * VEC);
      }
  }
}

template <int VEC, int MODE>
static int launch_gather_v(nb_ctx *ctx, float *out, const float *table, uint64_t table_pitch, const float *cache, uint64_t cache_pitch,
                           const uint32_t *cache_map, const float *const *shards, uint32_t n_shards, const uint32_t *ids,
                           uint32_t n_rows, const uint32_t *n_rows_dev, uint32_t F, uint64_t out_pitch, uint32_t *hit_count) {
  const uint32_t nvec = F / VEC, per_lane = (nvec + 31) / 32;
  const unsigned grid = nb_grid(n_rows, GATHER_THREADS / 32, 8);
#define NB_G(C) k_gather_rows<VEC, C, MODE><<<grid, GATHER_THREADS, 0, ctx->stream>>>(out, table, table_pitch, cache, cache_pitch, \
      cache_map, shards, n_shards, ids, n_rows, n_rows_dev, nvec, out_pitch, hit_count)
  if (per_lane <= 1) NB_G(1); else if (per_lane <= 2) NB_G(2); else if (per_lane <= 4) NB_G(4);
  else if (per_lane <= 5) NB_G(5); else if (per_lane <= 8) NB_G(8); else NB_G(10);
#undef NB_G
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
}

template <int MODE>
static int launch_gather(nb_ctx *ctx, float *out, const float *table, uint64_t table_pitch, const float *cache, uint64_t cache_pitch,
                         const uint32_t *cache_map, const float *const *shards, uint32_t n_shards, const uint32_t *ids,
                         uint32_t n_rows, uint32_t F, uint64_t out_pitch, uint32_t *hit_count, int vec,
                         const uint32_t *n_rows_dev = nullptr) {
  if (n_rows == 0) return NB_OK;
  if (vec == 4) return launch_gather_v<4, MODE>(ctx, out, table, table_pitch, cache, cache_pitch, cache_map, shards, n_shards, ids, n_rows, n_rows_dev, F, out_pitch, hit_count);
  if (vec == 2) return launch_gather_v<2, MODE>(ctx, out, table, table_pitch, cache, cache_pitch, cache_map, shards, n_shards, ids, n_rows, n_rows_dev, F, out_pitch, hit_count);
  return launch_gather_v<1, MODE>(ctx, out, table, table_pitch, cache, cache_pitch, cache_map, shards, n_shards, ids, n_rows, n_rows_dev, F, out_pitch, hit_count);
}

extern "C" {

int nb_gather_rows(nb_ctx *ctx, float *out, const float *table, const uint32_t *ids_dev, uint32_t n_rows,
                   uint32_t feature_size, uint32_t table_pitch, uint32_t out_pitch) {
  NB_REQUIRE(ctx && (n_rows == 0 || (out && table && ids_dev)), NB_ERR_ARG, "nb_gather_rows: NULL argument");
  NB_REQUIRE(feature_size > 0 && table_pitch >= feature_size && out_pitch >= feature_size, NB_ERR_ARG, "nb_gather_rows: bad pitch");
  NB_GUARD(ctx);
  int vec = nb_pick_vec(feature_size, table, table_pitch, out, out_pitch);
  return launch_gather<0>(ctx, out, table, table_pitch, nullptr, 0, nullptr, nullptr, 0, ids_dev, n_rows, feature_size, out_pitch, nullptr, vec);
}

int nb_gather_rows_dyn(nb_ctx *ctx, float *out, const float *table, const uint32_t *ids_dev, const uint32_t *n_rows_dev,
                       uint32_t max_rows, uint32_t feature_size, uint32_t table_pitch, uint32_t out_pitch) {
  NB_REQUIRE(ctx && out && table && ids_dev && n_rows_dev, NB_ERR_ARG, "nb_gather_rows_dyn: NULL argument");
  NB_REQUIRE(feature_size > 0 && table_pitch >= feature_size && out_pitch >= feature_size, NB_ERR_ARG, "nb_gather_rows_dyn: bad pitch");
  NB_GUARD(ctx);
  int vec = nb_pick_vec(feature_size, table, table_pitch, out, out_pitch);
  return launch_gather<0>(ctx, out, table, table_pitch, nullptr, 0, nullptr, nullptr, 0, ids_dev, max_rows, feature_size, out_pitch, nullptr, vec, n_rows_dev);
}

int nb_gather_rows_cached(nb_ctx *ctx, float *out, const float *cold_table, uint32_t cold_pitch, const float *cache_table,
                          uint32_t cache_pitch, const uint32_t *cache_node_hashmap_dev, const uint32_t *ids_dev,
                          uint32_t n_rows, uint32_t feature_size, uint32_t out_pitch, uint32_t *hit_count_dev_or_null) {
  NB_REQUIRE(ctx && (n_rows == 0 || (out && cold_table && cache_table && cache_node_hashmap_dev && ids_dev)), NB_ERR_ARG,
             "nb_gather_rows_cached: NULL argument");
  NB_REQUIRE(feature_size > 0 && cold_pitch >= feature_size && cache_pitch >= feature_size && out_pitch >= feature_size, NB_ERR_ARG, "bad pitch");
  NB_GUARD(ctx);
  int v1 = nb_pick_vec(feature_size, cold_table, cold_pitch, out, out_pitch);
  int v2 = nb_pick_vec(feature_size, cache_table, cache_pitch, out, out_pitch);
  return launch_gather<1>(ctx, out, cold_table, cold_pitch, cache_table, cache_pitch, cache_node_hashmap_dev, nullptr, 0, ids_dev,
                          n_rows, feature_size, out_pitch, hit_count_dev_or_null, v1 < v2 ? v1 : v2);
}

int nb_gather_labels(nb_ctx *ctx, int64_t *out, const int64_t *labels_dev, const uint32_t *ids_dev, uint32_t n) {
  NB_REQUIRE(ctx && (n == 0 || (out && labels_dev && ids_dev)), NB_ERR_ARG, "nb_gather_labels: NULL argument");
  NB_GUARD(ctx);
  if (n == 0) return NB_OK;
  k_gather_labels<<<nb_grid(n, 256, 4), 256, 0, ctx->stream>>>(out, labels_dev, ids_dev, n);
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
}

int nb_row_override2(nb_ctx *ctx, float *out_feature, float *out_embedding, const float *share_feature,
                     const float *share_embedding, const uint32_t *cache_map_dev, const uint32_t *cache_location_dev,
                     const uint32_t *destination_dev, uint32_t n_dst, uint32_t feature_size, uint32_t embedding_size,
                     uint32_t super_batch_id) {
  NB_REQUIRE(ctx && (n_dst == 0 || (out_feature && share_feature && cache_map_dev && cache_location_dev && destination_dev)),
             NB_ERR_ARG, "nb_row_override: NULL argument");
  NB_REQUIRE(feature_size > 0 && (!out_embedding || (share_embedding && embedding_size > 0)), NB_ERR_ARG, "nb_row_override: bad sizes");
  NB_GUARD(ctx);
  if (n_dst == 0) return NB_OK;
  int vec = nb_pick_vec(feature_size, out_feature, feature_size, share_feature, feature_size);
  if (out_embedding) {
    int v2 = nb_pick_vec(embedding_size, out_embedding, embedding_size, share_embedding, embedding_size);
    if (v2 < vec) vec = v2;
  }
  unsigned grid = nb_grid(n_dst, GATHER_THREADS / 32, 8);
#define NB_OVR(V)                                                                                                   \
  k_row_override<V><<<grid, GATHER_THREADS, 0, ctx->stream>>>(out_feature, share_feature, feature_size / V, feature_size, \
      out_embedding, share_embedding, out_embedding ? embedding_size / V : 0, embedding_size, cache_map_dev,          \
      cache_location_dev, destination_dev, n_dst, super_batch_id)
  if (vec == 4) NB_OVR(4); else if (vec == 2) NB_OVR(2); else NB_OVR(1);
#undef NB_OVR
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
}

int nb_row_override(nb_ctx *ctx, float *out, const float *share, const uint32_t *cache_map_dev,
                    const uint32_t *cache_location_dev, const uint32_t *destination_dev, uint32_t n_dst,
                    uint32_t feature_size, uint32_t super_batch_id) {
  return nb_row_override2(ctx, out, nullptr, share, nullptr, cache_map_dev, cache_location_dev, destination_dev, n_dst,
                          feature_size, 0, super_batch_id);
}

int nb_table_create(nb_ctx *ctx, uint32_t n_shards, const float *const *shard_ptrs, uint32_t feature_size,
                    uint32_t pitch, uint64_t n_rows_total, nb_table **out) {
  NB_REQUIRE(ctx && out && shard_ptrs && n_shards >= 1, NB_ERR_ARG, "nb_table_create: bad argument");
  NB_REQUIRE(feature_size > 0 && pitch >= feature_size, NB_ERR_ARG, "nb_table_create: bad pitch");
  NB_GUARD(ctx);
  nb_table *t = new nb_table();
  t->ctx = ctx; t->n_shards = n_shards; t->feature_size = feature_size; t->pitch = pitch; t->n_rows = n_rows_total;
  NB_CUDA(cudaMalloc(&t->shards_dev, sizeof(float *) * n_shards));
  NB_CUDA(cudaMemcpyAsync(t->shards_dev, shard_ptrs, sizeof(float *) * n_shards, cudaMemcpyHostToDevice, ctx->stream));
  NB_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = t;
  return NB_OK;
}

int nb_table_destroy(nb_table *t) {
  if (!t) return NB_OK;
  DeviceGuard guard(t->ctx->device);
  cudaFree(t->shards_dev);
  delete t;
  return NB_OK;
}

int nb_table_gather(nb_ctx *ctx, nb_table *t, float *out, const uint32_t *ids_dev, uint32_t n_rows, uint32_t out_pitch) {
  NB_REQUIRE(ctx && t && (n_rows == 0 || (out && ids_dev)), NB_ERR_ARG, "nb_table_gather: NULL argument");
  NB_REQUIRE(out_pitch >= t->feature_size, NB_ERR_ARG, "nb_table_gather: bad pitch");
  NB_GUARD(ctx);
  // shard bases come from cudaMalloc / IPC mappings: 256-byte aligned
  int vec = nb_pick_vec(t->feature_size, nullptr, t->pitch, out, out_pitch);
  return launch_gather<2>(ctx, out, nullptr, t->pitch, nullptr, 0, nullptr, t->shards_dev, t->n_shards, ids_dev, n_rows,
                          t->feature_size, out_pitch, nullptr, vec);
}

}  // extern "C"
