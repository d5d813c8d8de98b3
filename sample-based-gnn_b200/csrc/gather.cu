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
#include <stdlib.h>

#include "common.cuh"

constexpr int GATHER_THREADS = 256;
static int g_gather_narrow_rows = 4;  // "gather_narrow_rows": rows per warp iteration for rows of <= 32 vectors (1 = one row at a time)

// MODE 0: plain table. MODE 1: hot/cold (cache_map slot != -1 -> cache table). MODE 2: sharded table (v % n, v / n).
// MODE 3: three tiers -- hot rows (cache_map slot != -1) from a cache table that is either local (shards == NULL) or sharded over
// the GPUs (slot % n, slot / n), cold rows from the staged block of stage.cu (table = staged rows, cold_slot[i] = row of batch
// position i in that block).
// CHUNK = vectors per lane held in registers: a row of up to 32*CHUNK vectors is fetched with CHUNK independent
// requests per lane before the first store (602 floats = 301 float2 -> CHUNK 10, 2.4 KB in flight per warp);
// the next row's id is fetched one iteration ahead so the id -> row dependency is off the critical path.
template <int VEC, int CHUNK, int MODE>
__global__ void __launch_bounds__(GATHER_THREADS)
k_gather_rows(float *__restrict__ out, const float *__restrict__ table, uint64_t table_pitch, const float *__restrict__ cache,
              uint64_t cache_pitch, const uint32_t *__restrict__ cache_map, const float *const *__restrict__ shards,
              uint32_t n_shards, const uint32_t *__restrict__ ids, uint32_t n_rows, const uint32_t *__restrict__ n_rows_dev,
              uint32_t nvec, uint64_t out_pitch, uint32_t *hit_count, const uint32_t *__restrict__ cold_slot = nullptr) {
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
    } else if (MODE == 2) src = shards[v % n_shards] + (uint64_t)(v / n_shards) * table_pitch;
    else {
      const uint32_t slot = cache_map[v];
      if (slot == 0xffffffffu) src = table + (uint64_t)cold_slot[i] * table_pitch;
      else if (shards) src = shards[slot % n_shards] + (uint64_t)(slot / n_shards) * cache_pitch;
      else src = cache + (uint64_t)slot * cache_pitch;
    }
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

// Narrow rows (<= 32 vectors, e.g. F = 100 or 128): one vector per lane per row, so a warp that moves one row at a time has a
// single 16-byte request per lane in flight and the kernel is latency bound (~2.6 TB/s algorithmic at F = 100). Here a warp owns
// ROWS consecutive rows per iteration: ROWS ids in one broadcast load, ROWS independent row loads, then ROWS stores.
template <int VEC, int ROWS, int MODE>
__global__ void __launch_bounds__(GATHER_THREADS)
k_gather_rows_narrow(float *__restrict__ out, const float *__restrict__ table, uint64_t table_pitch, const float *__restrict__ cache,
                     uint64_t cache_pitch, const uint32_t *__restrict__ cache_map, const float *const *__restrict__ shards,
                     uint32_t n_shards, const uint32_t *__restrict__ ids, uint32_t n_rows, const uint32_t *__restrict__ n_rows_dev,
                     uint32_t nvec, uint64_t out_pitch, uint32_t *hit_count, const uint32_t *__restrict__ cold_slot) {
  const unsigned lane = lane_id();
  const unsigned warp = (blockIdx.x * GATHER_THREADS + threadIdx.x) >> 5;
  const unsigned warps = (gridDim.x * GATHER_THREADS) >> 5;
  if (n_rows_dev) n_rows = min(n_rows, *n_rows_dev);
  unsigned hits = 0;
  for (unsigned i0 = warp * ROWS; i0 < n_rows; i0 += warps * ROWS) {
    const float *src[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
      const unsigned i = i0 + r;
      src[r] = nullptr;
      if (i < n_rows) {
        const uint32_t v = ids[i];
        if (MODE == 0) src[r] = table + (uint64_t)v * table_pitch;
        else if (MODE == 1) {
          const uint32_t slot = cache_map[v];
          if (slot != 0xffffffffu) { src[r] = cache + (uint64_t)slot * cache_pitch; hits++; }
          else src[r] = table + (uint64_t)v * table_pitch;
        } else if (MODE == 2) src[r] = shards[v % n_shards] + (uint64_t)(v / n_shards) * table_pitch;
        else {
          const uint32_t slot = cache_map[v];
          if (slot == 0xffffffffu) src[r] = table + (uint64_t)cold_slot[i] * table_pitch;
          else if (shards) src[r] = shards[slot % n_shards] + (uint64_t)(slot / n_shards) * cache_pitch;
          else src[r] = cache + (uint64_t)slot * cache_pitch;
        }
      }
    }
    Vec<VEC> x[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; r++)
      if (src[r] && lane < nvec) x[r].load(src[r] + (uint64_t)lane * VEC);
#pragma unroll
    for (int r = 0; r < ROWS; r++)
      if (src[r] && lane < nvec) x[r].store(out + (uint64_t)(i0 + r) * out_pitch + (uint64_t)lane * VEC);
  }
  if (MODE == 1 && hit_count && lane == 0 && hits) atomicAdd(hit_count, hits);
}

// List-indirected gather, the reference's cached-load pair (zero_copy_feature_move_gpu_cache_kernel /
// gather_feature_from_gpu_cache_kernel, cuda/ntsCUDATransferKernel.cuh:154-183): for i in [0, n)
//   lid = local_idx[i];  v = ids[lid];  row = slot_map ? slot_map[v] : v;  out[lid,:] = table[row,:]
// local_idx / slot_map may live in mapped pinned host memory (the toolkits build both on the CPU). A warp owns ROWS list
// entries per iteration so that the three dependent index loads of ROWS rows overlap; rows move as VEC-wide vectors.
template <int VEC, int ROWS>
__global__ void __launch_bounds__(GATHER_THREADS)
k_gather_rows_indexed(float *__restrict__ out, uint64_t out_pitch, const float *__restrict__ table, uint64_t table_pitch,
                      const uint32_t *__restrict__ ids, const uint32_t *__restrict__ local_idx, const uint32_t *__restrict__ slot_map,
                      uint32_t n, uint32_t nvec) {
  const unsigned lane = lane_id();
  const unsigned warp = (blockIdx.x * GATHER_THREADS + threadIdx.x) >> 5;
  const unsigned warps = (gridDim.x * GATHER_THREADS) >> 5;
  for (unsigned i0 = warp * ROWS; i0 < n; i0 += warps * ROWS) {
    uint32_t lid[ROWS], row[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; r++) lid[r] = i0 + r < n ? local_idx[i0 + r] : 0u;
#pragma unroll
    for (int r = 0; r < ROWS; r++) row[r] = i0 + r < n ? ids[lid[r]] : 0u;
    if (slot_map) {
#pragma unroll
      for (int r = 0; r < ROWS; r++) row[r] = i0 + r < n ? slot_map[row[r]] : 0u;
    }
    for (unsigned k0 = 0; k0 < nvec; k0 += 32) {
      const unsigned k = k0 + lane;
      Vec<VEC> x[ROWS];
#pragma unroll
      for (int r = 0; r < ROWS; r++)
        if (i0 + r < n && k < nvec) x[r].load(table + (uint64_t)row[r] * table_pitch + (uint64_t)k * VEC);
#pragma unroll
      for (int r = 0; r < ROWS; r++)
        if (i0 + r < n && k < nvec) x[r].store(out + (uint64_t)lid[r] * out_pitch + (uint64_t)k * VEC);
    }
  }
}

// ---- TMA variant -------------------------------------------------------------------------------
// Rows whose byte length and addresses are 16-byte aligned (row pitches that are multiples of 4 floats, e.g. the
// 602-float Reddit row stored at pitch 608 = 19 x 128 B) move global -> shared -> global with bulk async copies
// (cp.async.bulk, SASS UBLKCP) and never touch registers. One thread owns one shared-memory row slot and its
// mbarrier; SLOTS rows (~2.4 KB each) are in flight per CTA, one CTA per SM, i.e. ~150 KB in flight per SM versus
// the <= 64 KB the register path can hold.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(128, 1)
k_gather_rows_tma(float *__restrict__ out, const float *__restrict__ table, uint64_t table_pitch, const float *__restrict__ cache,
                  uint64_t cache_pitch, const uint32_t *__restrict__ cache_map, const float *const *__restrict__ shards,
                  uint32_t n_shards, const uint32_t *__restrict__ ids, uint32_t n_rows, const uint32_t *__restrict__ n_rows_dev,
                  uint32_t copy_bytes, uint32_t slot_bytes, uint64_t out_pitch, int slots) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t *bars = (uint64_t *)smem;                  // [slots]
  uint8_t *rows = smem + ((slots * 8 + 127) & ~127);  // [slots][slot_bytes]
  if (n_rows_dev) n_rows = min(n_rows, *n_rows_dev);
  const int t = threadIdx.x;
  if (t >= slots) return;
  const uint32_t bar = smem_u32(&bars[t]);
  const uint32_t slot = smem_u32(rows + (size_t)t * slot_bytes);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const unsigned stride = gridDim.x * slots;
  unsigned i = blockIdx.x * slots + t;
  uint32_t v_next = i < n_rows ? ids[i] : 0;
  uint32_t phase = 0;
  for (; i < n_rows; i += stride) {
    const uint32_t v = v_next;
    if (i + stride < n_rows) v_next = ids[i + stride];
    const float *src;
    if (MODE == 0) src = table + (uint64_t)v * table_pitch;
    else if (MODE == 1) {
      const uint32_t s = cache_map[v];
      src = s != 0xffffffffu ? cache + (uint64_t)s * cache_pitch : table + (uint64_t)v * table_pitch;
    } else src = shards[v % n_shards] + (uint64_t)(v / n_shards) * table_pitch;
    // the previous store out of this slot must have finished reading it
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(copy_bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(slot), "l"(src), "r"(copy_bytes), "r"(bar) : "memory");
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(bar), "r"(phase) : "memory");
    }
    phase ^= 1;
    float *dst = out + (uint64_t)i * out_pitch;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(slot), "r"(copy_bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// returns the number of bytes a bulk row copy may move (0 = layout not eligible)
static uint32_t tma_row_bytes(uint32_t F, const void *a, uint64_t pitch_a, const void *b, uint64_t pitch_b, const void *c = nullptr, uint64_t pitch_c = 4) {
  uint32_t bytes = (F * 4 + 31) & ~31u;   // whole 32-byte sectors when every pitch has the room (no partial-sector write per row)
  if (pitch_a * 4 < bytes || pitch_b * 4 < bytes || (c && pitch_c * 4 < bytes)) bytes = (F * 4 + 15) & ~15u;
  if (pitch_a % 4 || pitch_b % 4 || pitch_c % 4) return 0;
  if (pitch_a * 4 < bytes || pitch_b * 4 < bytes || (c && pitch_c * 4 < bytes)) return 0;
  if ((uintptr_t)a % 16 || (uintptr_t)b % 16 || (uintptr_t)c % 16) return 0;
  return bytes;
}

static int g_table_gather_tma = 0;   // "table_gather_tma": sharded-table gather through TMA bulk copies (A/B; tools/shard_bench.py)
static int g_gather_variant = -1;  // NB_GATHER_VARIANT: 0 = registers (LDG/STG), 1 = TMA bulk when eligible (default)
static int gather_variant() {
  if (g_gather_variant < 0) {
    const char *e = getenv("NB_GATHER_VARIANT");
    g_gather_variant = e ? atoi(e) : 1;
  }
  return g_gather_variant;
}

template <int MODE>
static int launch_gather_tma(nb_ctx *ctx, float *out, const float *table, uint64_t table_pitch, const float *cache, uint64_t cache_pitch,
                             const uint32_t *cache_map, const float *const *shards, uint32_t n_shards, const uint32_t *ids,
                             uint32_t n_rows, const uint32_t *n_rows_dev, uint32_t copy_bytes, uint64_t out_pitch) {
  const uint32_t slot_bytes = (copy_bytes + 127) & ~127u;
  int slots = (int)((200 * 1024) / slot_bytes);
  if (slots > 128) slots = 128;
  if (slots < 1) return NB_ERR_UNSUPPORTED;
  const size_t smem = ((slots * 8 + 127) & ~127) + (size_t)slots * slot_bytes;
  static bool attr_set[3] = {false, false, false};
  if (!attr_set[MODE]) {
    NB_CUDA(cudaFuncSetAttribute(k_gather_rows_tma<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_set[MODE] = true;
  }
  unsigned grid = (n_rows + slots - 1) / slots;
  if (grid > (unsigned)ctx->sm_count) grid = ctx->sm_count;
  k_gather_rows_tma<MODE><<<grid, 128, smem, ctx->stream>>>(out, table, table_pitch, cache, cache_pitch, cache_map, shards, n_shards, ids,
                                                           n_rows, n_rows_dev, copy_bytes, slot_bytes, out_pitch, slots);
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
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
                           uint32_t n_rows, const uint32_t *n_rows_dev, uint32_t F, uint64_t out_pitch, uint32_t *hit_count,
                           const uint32_t *cold_slot = nullptr) {
  const uint32_t nvec = F / VEC, per_lane = (nvec + 31) / 32;
  if (per_lane <= 1 && g_gather_narrow_rows > 1) {
    constexpr int ROWS = 4;
    const unsigned grid = nb_grid(n_rows, GATHER_THREADS / 32 * ROWS, 8);
    k_gather_rows_narrow<VEC, ROWS, MODE><<<grid, GATHER_THREADS, 0, ctx->stream>>>(out, table, table_pitch, cache, cache_pitch, cache_map,
        shards, n_shards, ids, n_rows, n_rows_dev, nvec, out_pitch, hit_count, cold_slot);
    NB_LAUNCH_CHECK(ctx);
    return NB_OK;
  }
  const unsigned grid = nb_grid(n_rows, GATHER_THREADS / 32, 8);
#define NB_G(C) k_gather_rows<VEC, C, MODE><<<grid, GATHER_THREADS, 0, ctx->stream>>>(out, table, table_pitch, cache, cache_pitch, \
      cache_map, shards, n_shards, ids, n_rows, n_rows_dev, nvec, out_pitch, hit_count, cold_slot)
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
                         const uint32_t *n_rows_dev = nullptr, const uint32_t *cold_slot = nullptr) {
  if (n_rows == 0) return NB_OK;
  if (vec == 4) return launch_gather_v<4, MODE>(ctx, out, table, table_pitch, cache, cache_pitch, cache_map, shards, n_shards, ids, n_rows, n_rows_dev, F, out_pitch, hit_count, cold_slot);
  if (vec == 2) return launch_gather_v<2, MODE>(ctx, out, table, table_pitch, cache, cache_pitch, cache_map, shards, n_shards, ids, n_rows, n_rows_dev, F, out_pitch, hit_count, cold_slot);
  return launch_gather_v<1, MODE>(ctx, out, table, table_pitch, cache, cache_pitch, cache_map, shards, n_shards, ids, n_rows, n_rows_dev, F, out_pitch, hit_count, cold_slot);
}

// stage.cu's merge: hot rows from the (local or sharded) cache table, cold rows from the staged block
int nb_launch_gather_tiered(nb_ctx *ctx, float *out, uint64_t out_pitch, const float *cache, uint64_t cache_pitch, const nb_table *hot,
                            const uint32_t *cache_map, const uint32_t *ids, uint32_t n_rows, const float *staged, uint64_t staged_pitch,
                            const uint32_t *cold_slot, uint32_t F) {
  if (hot) cache_pitch = hot->pitch;
  int vec = nb_pick_vec(F, staged, staged_pitch, out, out_pitch);
  const int v2 = nb_pick_vec(F, hot ? nullptr : cache, cache_pitch, out, out_pitch);
  if (v2 < vec) vec = v2;
  return launch_gather<3>(ctx, out, staged, staged_pitch, cache, cache_pitch, cache_map, hot ? hot->shards_dev : nullptr,
                          hot ? hot->n_shards : 0, ids, n_rows, F, out_pitch, nullptr, vec, nullptr, cold_slot);
}

extern "C" {

int nb_set_option(const char *name, int value) {
  NB_REQUIRE(name, NB_ERR_ARG, "nb_set_option: NULL name");
  if (!strcmp(name, "gather_variant")) { g_gather_variant = value; return NB_OK; }
  if (!strcmp(name, "table_gather_tma")) { g_table_gather_tma = value; return NB_OK; }
  if (!strcmp(name, "gather_narrow_rows")) { g_gather_narrow_rows = value; return NB_OK; }
  if (!strcmp(name, "agg_blocks_per_sm")) { nb_agg_set_option(0, value); return NB_OK; }
  if (!strcmp(name, "agg_persistent")) { nb_agg_set_option(1, value); return NB_OK; }
  if (!strcmp(name, "agg_long_rows")) { nb_agg_set_option(2, value); return NB_OK; }
  if (!strcmp(name, "agg_pipe_wide")) { nb_agg_set_option(3, value); return NB_OK; }
  if (!strcmp(name, "agg_short_rows")) { nb_agg_set_option(4, value); return NB_OK; }   // CSR backward: 4 rows per warp in flight (default 0: no gain measured)
  if (!strcmp(name, "agg_deep_small")) { nb_agg_set_option(5, value); return NB_OK; }   // small launches: 16 / 8 entries in flight (default 1)
  if (!strcmp(name, "sampler_fused")) { nb_sampler_set_fused(value); return NB_OK; }   // read when a sampler is created
  if (!strcmp(name, "sampler_two_level")) { nb_sampler_set_two_level(value); return NB_OK; }
  if (!strcmp(name, "peer_push_side_stream")) { nb_peer_set_push_side(value); return NB_OK; }   // read by nb_peer_comm_create
  if (!strcmp(name, "sampler_block_threads")) { nb_sampler_set_block(value); return NB_OK; }   // read when a sampler's graph is captured
  if (!strcmp(name, "sampler_blocks_per_sm")) { nb_sampler_set_bps(value); return NB_OK; }   // read when a sampler's graph is captured
  if (!strcmp(name, "sampler_capture_priority")) { nb_sampler_set_capture_prio(value); return NB_OK; }
  if (!strcmp(name, "sampler_csr_branch")) { nb_sampler_set_csr_branch(value); return NB_OK; }
  if (!strcmp(name, "sampler_tail")) { nb_sampler_set_tail(value); return NB_OK; }   // read when a sampler's graph is captured
  if (!strcmp(name, "gather_keep_min_uses")) { nb_sampler_set_keep_min(value); return NB_OK; }   // read when a sampler's graph is captured
  if (!strcmp(name, "trace")) { nb_trace_set_level(value); return NB_OK; }
  if (!strcmp(name, "mirror_host_tables")) { nb_mirror_host_enable(value, 0); return NB_OK; }
  if (!strcmp(name, "mirror_host_adjacency")) { nb_mirror_host_enable(value, 1); return NB_OK; }
  nb_set_error("nb_set_option: unknown option %s", name);
  return NB_ERR_ARG;
}

int nb_gather_rows(nb_ctx *ctx, float *out, const float *table, const uint32_t *ids_dev, uint32_t n_rows,
                   uint32_t feature_size, uint32_t table_pitch, uint32_t out_pitch) {
  NB_REQUIRE(ctx && (n_rows == 0 || (out && table && ids_dev)), NB_ERR_ARG, "nb_gather_rows: NULL argument");
  NB_REQUIRE(feature_size > 0 && table_pitch >= feature_size && out_pitch >= feature_size, NB_ERR_ARG, "nb_gather_rows: bad pitch");
  NB_GUARD(ctx);
  if (n_rows == 0) return NB_OK;
  table = (const float *)nb_mirror_host(ctx, table);
  if (gather_variant() == 2) {   // tile::gather4 through tensor maps (gather4.cu): A/B variant, falls through when the shape is not eligible
    const int rc = nb_gather4_launch(ctx, out, out_pitch, table, table_pitch, ids_dev, n_rows, feature_size);
    if (rc != NB_ERR_UNSUPPORTED) return rc;
  }
  // rows of <= 128 floats: the register path with 4 rows per warp beats bulk copies of 400-512 byte rows (tools/gather_bench.py)
  const uint32_t tb = gather_variant() >= 1 && feature_size > 128 ? tma_row_bytes(feature_size, table, table_pitch, out, out_pitch) : 0;
  if (tb) return launch_gather_tma<0>(ctx, out, table, table_pitch, nullptr, 0, nullptr, nullptr, 0, ids_dev, n_rows, nullptr, tb, out_pitch);
  uint32_t fe = feature_size;
  int vec = nb_pick_vec(feature_size, table, table_pitch, out, out_pitch, &fe);
  return launch_gather<0>(ctx, out, table, table_pitch, nullptr, 0, nullptr, nullptr, 0, ids_dev, n_rows, fe, out_pitch, nullptr, vec);
}

int nb_gather_rows_dyn(nb_ctx *ctx, float *out, const float *table, const uint32_t *ids_dev, const uint32_t *n_rows_dev,
                       uint32_t max_rows, uint32_t feature_size, uint32_t table_pitch, uint32_t out_pitch) {
  NB_REQUIRE(ctx && (max_rows == 0 || (out && table && ids_dev)), NB_ERR_ARG, "nb_gather_rows_dyn: NULL argument");
  NB_REQUIRE(feature_size > 0 && table_pitch >= feature_size && out_pitch >= feature_size, NB_ERR_ARG, "nb_gather_rows_dyn: bad pitch");
  NB_GUARD(ctx);
  if (max_rows == 0) return NB_OK;
  table = (const float *)nb_mirror_host(ctx, table);
  const uint32_t tb = gather_variant() == 1 && feature_size > 128 ? tma_row_bytes(feature_size, table, table_pitch, out, out_pitch) : 0;
  if (tb) return launch_gather_tma<0>(ctx, out, table, table_pitch, nullptr, 0, nullptr, nullptr, 0, ids_dev, max_rows, n_rows_dev, tb, out_pitch);
  uint32_t fe = feature_size;
  int vec = nb_pick_vec(feature_size, table, table_pitch, out, out_pitch, &fe);
  return launch_gather<0>(ctx, out, table, table_pitch, nullptr, 0, nullptr, nullptr, 0, ids_dev, max_rows, fe, out_pitch, nullptr, vec, n_rows_dev);
}

int nb_gather_rows_cached(nb_ctx *ctx, float *out, const float *cold_table, uint32_t cold_pitch, const float *cache_table,
                          uint32_t cache_pitch, const uint32_t *cache_node_hashmap_dev, const uint32_t *ids_dev,
                          uint32_t n_rows, uint32_t feature_size, uint32_t out_pitch, uint32_t *hit_count_dev_or_null) {
  NB_REQUIRE(ctx && (n_rows == 0 || (out && cold_table && cache_table && cache_node_hashmap_dev && ids_dev)), NB_ERR_ARG,
             "nb_gather_rows_cached: NULL argument");
  NB_REQUIRE(feature_size > 0 && cold_pitch >= feature_size && cache_pitch >= feature_size && out_pitch >= feature_size, NB_ERR_ARG, "bad pitch");
  NB_GUARD(ctx);
  cold_table = (const float *)nb_mirror_host(ctx, cold_table);
  int v1 = nb_pick_vec(feature_size, cold_table, cold_pitch, out, out_pitch);
  int v2 = nb_pick_vec(feature_size, cache_table, cache_pitch, out, out_pitch);
  return launch_gather<1>(ctx, out, cold_table, cold_pitch, cache_table, cache_pitch, cache_node_hashmap_dev, nullptr, 0, ids_dev,
                          n_rows, feature_size, out_pitch, hit_count_dev_or_null, v1 < v2 ? v1 : v2);
}

int nb_gather_rows_indexed(nb_ctx *ctx, float *out, uint32_t out_pitch, const float *table, uint32_t table_pitch,
                           const uint32_t *ids_dev, const uint32_t *local_idx, const uint32_t *slot_map_or_null, uint32_t n,
                           uint32_t feature_size) {
  NB_REQUIRE(ctx && (n == 0 || (out && table && ids_dev && local_idx)), NB_ERR_ARG, "nb_gather_rows_indexed: NULL argument");
  NB_REQUIRE(feature_size > 0 && table_pitch >= feature_size && out_pitch >= feature_size, NB_ERR_ARG, "nb_gather_rows_indexed: bad pitch");
  NB_GUARD(ctx);
  if (n == 0) return NB_OK;
  table = (const float *)nb_mirror_host(ctx, table);
  uint32_t fe = feature_size;
  const int vec = nb_pick_vec(feature_size, table, table_pitch, out, out_pitch, &fe);
  const unsigned grid = nb_grid(n, GATHER_THREADS / 32 * 4, 8);
  if (vec == 4) k_gather_rows_indexed<4, 4><<<grid, GATHER_THREADS, 0, ctx->stream>>>(out, out_pitch, table, table_pitch, ids_dev, local_idx, slot_map_or_null, n, fe / 4);
  else if (vec == 2) k_gather_rows_indexed<2, 4><<<grid, GATHER_THREADS, 0, ctx->stream>>>(out, out_pitch, table, table_pitch, ids_dev, local_idx, slot_map_or_null, n, fe / 2);
  else k_gather_rows_indexed<1, 4><<<grid, GATHER_THREADS, 0, ctx->stream>>>(out, out_pitch, table, table_pitch, ids_dev, local_idx, slot_map_or_null, n, fe);
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
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
  if (g_table_gather_tma && t->n_shards > 1) {   // "table_gather_tma": rows of peer shards as TMA bulk copies (thousands in flight per SM)
    const uint32_t tb = tma_row_bytes(t->feature_size, nullptr, t->pitch, out, out_pitch);
    if (tb) return launch_gather_tma<2>(ctx, out, nullptr, t->pitch, nullptr, 0, nullptr, t->shards_dev, t->n_shards, ids_dev, n_rows, nullptr, tb, out_pitch);
  }
  int vec = nb_pick_vec(t->feature_size, nullptr, t->pitch, out, out_pitch);
  return launch_gather<2>(ctx, out, nullptr, t->pitch, nullptr, 0, nullptr, t->shards_dev, t->n_shards, ids_dev, n_rows,
                          t->feature_size, out_pitch, nullptr, vec);
}

}  // extern "C"
