// common.cuh -- shared device/host helpers for libnts_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/nts_b200.h"

#define NB_SM_COUNT 148  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// ---- error plumbing ---------------------------------------------------------------------------
void nb_set_error(const char *fmt, ...);
#define NB_CUDA(call)                                                                              \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      nb_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));          \
      return NB_ERR_CUDA;                                                                          \
    }                                                                                              \
  } while (0)
#define NB_REQUIRE(cond, code, ...)                                                                \
  do {                                                                                             \
    if (!(cond)) {                                                                                 \
      nb_set_error(__VA_ARGS__);                                                                   \
      return (code);                                                                               \
    }                                                                                              \
  } while (0)
#define NB_LAUNCH_CHECK(ctx)                                                                       \
  do {                                                                                             \
    (ctx)->launches++;                                                                             \
    cudaError_t e__ = cudaPeekAtLastError();                                                       \
    if (e__ != cudaSuccess) {                                                                      \
      nb_set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__));      \
      return NB_ERR_CUDA;                                                                          \
    }                                                                                              \
  } while (0)

struct nb_ctx {
  int device;
  cudaStream_t stream;
  bool own_stream;
  uint64_t launches;
  int sm_count;
  void *scratch;         // grow-only device scratch (replaces Cuda_Stream::cuda_buffer, cuda/ntsCUDA.hpp:195-196)
  size_t scratch_bytes;
};
int nb_ctx_scratch(nb_ctx *ctx, size_t bytes, void **out);
// segment reduction out[r,:] = sum_j w[j] in[idx[j],:] (+ e1[r] va + e2[r] vb); aggregate.cu
int nb_run_segment(nb_ctx *ctx, bool push, const float *in, float *out, const float *w, const uint32_t *idx, const uint32_t *offsets,
                   uint32_t n_rows, uint32_t F, const uint32_t *n_rows_dev, uint64_t in_pitch, uint64_t out_pitch, const float *e1,
                   const float *e2, const float *va, const float *vb, bool packed_index = false, int shape = 0);
constexpr int NB_SEG_SHORT_ROWS = 1;   // shape hint: rows are the CSR rows of a sampled layer (~1-2 entries each)
int nb_run_segment_gat(nb_ctx *ctx, const float *dout, float *dh, const uint32_t *column_indices, const uint32_t *row_offset, uint32_t n_src,
                       uint32_t F, const uint32_t *c2c, const float *alpha, const float *ds, const float *dsum, const uint32_t *src_to_dst,
                       const float *va, const float *vb, float *rs_out, float *dd_out);
// (optionally sharded) HBM feature table; gather.cu
struct nb_table {
  nb_ctx *ctx;
  uint32_t n_shards, feature_size, pitch;
  uint64_t n_rows;
  const float **shards_dev;  // device array of n_shards row-base pointers
};
int nb_launch_gather_tiered(nb_ctx *ctx, float *out, uint64_t out_pitch, const float *cache, uint64_t cache_pitch, const nb_table *hot,
                            const uint32_t *cache_map, const uint32_t *ids, uint32_t n_rows, const float *staged, uint64_t staged_pitch,
                            const uint32_t *cold_slot, uint32_t F);
const void *nb_mirror_host(nb_ctx *ctx, const void *p, int is_adjacency = 0);  // HBM copy of a mapped-host allocation (or p itself)
void nb_mirror_host_enable(int on, int adjacency);  // stream-ordered reuse; grows with cudaMalloc

struct DeviceGuard {
  int prev;
  bool ok;
  explicit DeviceGuard(int dev) : prev(-1), ok(true) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
// Per-entry-point wall-clock accounting (NB_TRACE=1: host time inside each nb_* call; NB_TRACE=2: the call's stream is
// synchronised before the clock stops, so the time includes the GPU work -- serialising, diagnostic only). A table is printed
// to stderr at exit, or on demand by nb_trace_dump(). Replaces the reference's manual get_time() accumulators
// (core/ntsFastSampler.hpp:30-37, cuda/ntsCUDA.hpp:180-198 cpu_inclusiveTime / inclusiveTime).
struct NbTraceScope {
  const char *name;
  const nb_ctx *ctx;   // the stream is read when the scope closes: nb_ctx_set_stream replaces (and destroys) it inside the call
  uint64_t t0;
  NbTraceScope(const char *fn, const nb_ctx *c);
  ~NbTraceScope();
};
void nb_trace_set_level(int level);
bool nb_trace_on();
uint64_t nb_trace_now_ns();
void nb_trace_add(const char *name, uint64_t ns);  // name must be a string literal (the table is keyed by its address)
void nb_sampler_set_two_level(int mode);
void nb_sampler_set_keep_min(int n);
// TMA tensor-map gather (tile::gather4), gather4.cu; NB_ERR_UNSUPPORTED = shape not eligible, caller falls back
int nb_gather4_launch(nb_ctx *ctx, float *out, uint64_t out_pitch, const float *table, uint64_t table_pitch, const uint32_t *ids_dev,
                      uint32_t n_rows, uint32_t F);
void nb_sampler_set_tail(int v);
void nb_sampler_set_csr_branch(int v);
void nb_sampler_set_block(int v);
void nb_sampler_set_bps(int v);
void nb_sampler_set_capture_prio(int v);
void nb_peer_set_push_side(int v);
void nb_sampler_set_fused(int on);   // sample.cu: small-shape sampler path on/off for samplers created afterwards
void nb_agg_set_option(int which, int value);  // 0: resident blocks per SM of the segment reduction (1..8), 1: persistent grid on/off, 2: block-per-row path for long segments on/off
#define NB_GUARD(ctx)                                                                              \
  DeviceGuard guard__((ctx)->device);                                                              \
  NB_REQUIRE(guard__.ok, NB_ERR_CUDA, "cudaSetDevice(%d) failed (no CUDA device?)", (ctx)->device); \
  NbTraceScope trace__(__func__, (ctx))

static inline unsigned nb_grid(uint64_t work_items, unsigned items_per_block, unsigned max_blocks_per_sm = 8) {
  uint64_t need = (work_items + items_per_block - 1) / items_per_block;
  uint64_t cap = (uint64_t)NB_SM_COUNT * max_blocks_per_sm;
  if (need < 1) need = 1;
  return (unsigned)(need < cap ? need : cap);
}

// ---- device helpers ---------------------------------------------------------------------------
#ifdef __CUDACC__
#define FULL_MASK 0xffffffffu

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

// 128-/64-/32-bit read-only loads that do not pollute L1 (rows are touched once per kernel)
__device__ __forceinline__ float4 ldg_stream4(const float *p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float2 ldg_stream2(const float *p) {
  float2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ float ldg_stream1(const float *p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
// streaming stores (written once, read by a later kernel through L2)
__device__ __forceinline__ void stg_stream4(float *p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void stg_stream2(float *p, float2 v) {
  asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void stg_stream1(float *p, float v) {
  asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// L2 eviction-priority hints (createpolicy): rows that will be read again soon are kept (evict_last), rows read once make room
// first (evict_first). Used by the gather-fused aggregation, where the sampler knows how often a batch uses every source row.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float4 ldg_hint4(const float *p, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float2 ldg_hint2(const float *p, uint64_t pol) {
  float2 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float ldg_hint1(const float *p, uint64_t pol) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  return v;
}

template <int VEC> struct Vec;
template <> struct Vec<4> {
  float4 v;
  __device__ __forceinline__ void load(const float *p) { v = ldg_stream4(p); }
  __device__ __forceinline__ void load_hint(const float *p, uint64_t pol) { v = ldg_hint4(p, pol); }
  __device__ __forceinline__ void load_cached(const float *p) { v = __ldg(reinterpret_cast<const float4 *>(p)); }  // L1-allocating: data every warp re-reads
  __device__ __forceinline__ void store(float *p) const { stg_stream4(p, v); }
  __device__ __forceinline__ void zero() { v = make_float4(0.f, 0.f, 0.f, 0.f); }
  // out = out + in*w : multiply, then add, each rounded (core/ntsBaseOp.hpp:546-562 spells it mul+add)
  __device__ __forceinline__ void axpy(const Vec<4> &in, float w) {
    v.x = __fadd_rn(v.x, __fmul_rn(in.v.x, w)); v.y = __fadd_rn(v.y, __fmul_rn(in.v.y, w));
    v.z = __fadd_rn(v.z, __fmul_rn(in.v.z, w)); v.w = __fadd_rn(v.w, __fmul_rn(in.v.w, w));
  }
  __device__ __forceinline__ float dot(const Vec<4> &o) const { return v.x * o.v.x + v.y * o.v.y + v.z * o.v.z + v.w * o.v.w; }
};
template <> struct Vec<2> {
  float2 v;
  __device__ __forceinline__ void load(const float *p) { v = ldg_stream2(p); }
  __device__ __forceinline__ void load_hint(const float *p, uint64_t pol) { v = ldg_hint2(p, pol); }
  __device__ __forceinline__ void load_cached(const float *p) { v = __ldg(reinterpret_cast<const float2 *>(p)); }
  __device__ __forceinline__ void store(float *p) const { stg_stream2(p, v); }
  __device__ __forceinline__ void zero() { v = make_float2(0.f, 0.f); }
  __device__ __forceinline__ void axpy(const Vec<2> &in, float w) {
    v.x = __fadd_rn(v.x, __fmul_rn(in.v.x, w)); v.y = __fadd_rn(v.y, __fmul_rn(in.v.y, w));
  }
  __device__ __forceinline__ float dot(const Vec<2> &o) const { return v.x * o.v.x + v.y * o.v.y; }
};
template <> struct Vec<1> {
  float v;
  __device__ __forceinline__ void load(const float *p) { v = ldg_stream1(p); }
  __device__ __forceinline__ void load_hint(const float *p, uint64_t pol) { v = ldg_hint1(p, pol); }
  __device__ __forceinline__ void load_cached(const float *p) { v = __ldg(p); }
  __device__ __forceinline__ void store(float *p) const { stg_stream1(p, v); }
  __device__ __forceinline__ void zero() { v = 0.f; }
  __device__ __forceinline__ void axpy(const Vec<1> &in, float w) { v = __fadd_rn(v, __fmul_rn(in.v, w)); }
  __device__ __forceinline__ float dot(const Vec<1> &o) const { return v * o.v; }
};

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL_MASK, x, o);
  return x;
}
__device__ __forceinline__ float warp_max(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(FULL_MASK, x, o));
  return x;
}
__device__ __forceinline__ unsigned warp_sum_u32(unsigned x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL_MASK, x, o);
  return x;
}

// ---- Philox4x32-10 (counter based; Salmon et al., SC'11) ---------------------------------------
struct Philox {
  uint32_t k0, k1;
  __device__ __forceinline__ Philox(uint64_t key) : k0((uint32_t)key), k1((uint32_t)(key >> 32)) {}
  __device__ __forceinline__ uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; r++) {
      uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      uint32_t n0 = hi1 ^ c1 ^ a, n1 = lo1, n2 = hi0 ^ c3 ^ b, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};
#endif  // __CUDACC__

// Vector width (4, 2 or 1 floats) usable for rows of `feature_size` floats at the given pitches / base pointers.
// *f_eff is the row length the kernel should process: when the rows are padded (pitch >= feature_size rounded
// up to the vector width, pitch a multiple of it) the pad columns are carried along so that a 602-float row at
// pitch 608 still moves as 128-bit vectors. Pad columns only ever flow into pad columns.
static inline int nb_pick_vec(uint32_t feature_size, const void *a, uint64_t pitch_a, const void *b, uint64_t pitch_b,
                              uint32_t *f_eff = nullptr) {
  uintptr_t pa = (uintptr_t)a, pb = (uintptr_t)b;
  for (int vec = 4; vec >= 2; vec >>= 1) {
    const uint32_t up = (feature_size + vec - 1) / vec * vec;
    if (pitch_a % vec == 0 && pitch_b % vec == 0 && pitch_a >= up && pitch_b >= up && pa % (4 * vec) == 0 && pb % (4 * vec) == 0) {
      if (f_eff) {
        // whole 32-byte sectors when the padding allows: a row that stops 16 bytes short of a sector boundary costs a partial-sector
        // write per row (602 floats at pitch 608: copying 604 runs 8 % slower than copying 608)
        const uint32_t up8 = (feature_size + 7) / 8 * 8;
        *f_eff = (pitch_a >= up8 && pitch_b >= up8) ? up8 : up;
      } else if (up != feature_size) continue;
      return vec;
    }
  }
  if (f_eff) *f_eff = feature_size;
  return 1;
}
