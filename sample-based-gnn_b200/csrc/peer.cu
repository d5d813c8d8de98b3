// peer.cu -- one-shot sum all-reduce of a small buffer over NVLink peer memory (one kernel, no NCCL).
//
// Replaces Parameter::reduce_multi_gpu_gradient -> NCCL_Communicator::AllReduce per tensor (core/NtsScheduler.hpp:830-836,
// cuda/ntsCUDAGraphOP.cu:173-200): the only exchange of the data-parallel path, ~330 KB of dense weight gradients per step --
// latency bound. Every rank owns one peer-shareable block (nb_vmm_alloc) mapped by all ranks:
//     [flags: world x PEER_CTAS u32][slot 0][slot 1]
// and one kernel of PEER_CTAS blocks does the whole exchange; block b owns chunk b of the buffer on every rank:
//   1. copy chunk b of the input into this rank's slot (seq & 1), __threadfence_system
//   2. store seq into flag[rank][b] of EVERY peer (remote 4-byte stores over NVLink)
//   3. wait until the local flag[p][b] of every peer p reached seq (local polling, bounded: a rank that never arrives raises an
//      error flag after ~20 s instead of hanging the GPU)
//   4. out[i] = slot_0[i] + slot_1[i] + ... in rank order, read straight from the peers' slots (volatile 128-bit loads): the same
//      order on every rank, so the result is bit-identical everywhere and run to run.
// Chunks never depend on each other, so there is no grid-wide barrier. Two slots: a rank can start exchange s+1 while a slow peer
// still reads slot s; it cannot start s+2 before every peer signalled s+1, i.e. finished reading s.
#include "common.cuh"

constexpr int PEER_CTAS = 16, PEER_THREADS = 256, PEER_MAX_WORLD = 16;
constexpr size_t PEER_FLAG_BYTES = 4096;   // world x PEER_CTAS u32, padded

struct PeerParams {
  uint8_t *base[PEER_MAX_WORLD];   // every rank's block as mapped on THIS device
  uint32_t rank, world;
  uint64_t slot_bytes;
};

struct nb_peer_comm {
  nb_ctx *ctx;
  PeerParams p;
  uint32_t seq;
  uint64_t max_floats;
  uint32_t *err_dev;
};

__device__ __forceinline__ float4 ld_volatile4(const float *p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(PEER_THREADS)
k_peer_allreduce(PeerParams c, float *__restrict__ inout, uint32_t n, uint32_t seq, uint32_t *__restrict__ err) {
  __shared__ int s_timeout;
  const uint32_t n4 = (n + 3) / 4;                                     // 16-byte units; slots are zero-padded to a multiple of 4 floats
  const uint32_t per = (n4 + gridDim.x - 1) / gridDim.x;
  const uint32_t lo = blockIdx.x * per, hi = min(n4, lo + per);
  float *my_slot = reinterpret_cast<float *>(c.base[c.rank] + PEER_FLAG_BYTES + (uint64_t)(seq & 1u) * c.slot_bytes);
  if (threadIdx.x == 0) s_timeout = 0;
  for (uint32_t i = lo + threadIdx.x; i < hi; i += PEER_THREADS) {
    float4 v;
    if (4 * i + 3 < n) v = *reinterpret_cast<const float4 *>(inout + 4 * (uint64_t)i);
    else {
      v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (4 * i < n) v.x = inout[4 * (uint64_t)i];
      if (4 * i + 1 < n) v.y = inout[4 * (uint64_t)i + 1];
      if (4 * i + 2 < n) v.z = inout[4 * (uint64_t)i + 2];
    }
    *reinterpret_cast<float4 *>(my_slot + 4 * (uint64_t)i) = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < c.world) {
    const uint32_t p = threadIdx.x;
    // 2. tell peer p that chunk b of this rank is in place
    volatile uint32_t *remote = reinterpret_cast<volatile uint32_t *>(c.base[p]) + c.rank * PEER_CTAS + blockIdx.x;
    *remote = seq;
    // 3. wait for peer p's chunk b
    volatile uint32_t *local = reinterpret_cast<volatile uint32_t *>(c.base[c.rank]) + p * PEER_CTAS + blockIdx.x;
    const long long t0 = clock64();
    while ((int)(*local - seq) < 0) {
      if (clock64() - t0 > 40000000000ll) { s_timeout = 1; break; }   // ~20 s at 2 GHz: a peer never arrived
    }
    __threadfence_system();
  }
  __syncthreads();
  if (s_timeout) {
    if (threadIdx.x == 0) atomicExch(err, 1u);
    return;
  }
  const uint64_t slot_off = PEER_FLAG_BYTES + (uint64_t)(seq & 1u) * c.slot_bytes;
  for (uint32_t i = lo + threadIdx.x; i < hi; i += PEER_THREADS) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint32_t p = 0; p < c.world; p++) {
      const float4 v = ld_volatile4(reinterpret_cast<const float *>(c.base[p] + slot_off) + 4 * (uint64_t)i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (4 * i + 3 < n) *reinterpret_cast<float4 *>(inout + 4 * (uint64_t)i) = acc;
    else {
      if (4 * i < n) inout[4 * (uint64_t)i] = acc.x;
      if (4 * i + 1 < n) inout[4 * (uint64_t)i + 1] = acc.y;
      if (4 * i + 2 < n) inout[4 * (uint64_t)i + 2] = acc.z;
    }
  }
}

extern "C" {

size_t nb_peer_comm_block_bytes(uint64_t max_floats) {
  const uint64_t slot = ((max_floats + 3) / 4 * 16 + 255) & ~255ull;
  return PEER_FLAG_BYTES + 2 * slot;
}

int nb_peer_comm_create(nb_ctx *ctx, uint32_t rank, uint32_t world, uint64_t max_floats, void *const *blocks, nb_peer_comm **out) {
  NB_REQUIRE(ctx && out && blocks && world >= 1 && world <= PEER_MAX_WORLD && rank < world && max_floats > 0, NB_ERR_ARG, "nb_peer_comm_create: bad argument");
  NB_REQUIRE((size_t)world * PEER_CTAS * 4 <= PEER_FLAG_BYTES, NB_ERR_ARG, "nb_peer_comm_create: world too large");
  NB_GUARD(ctx);
  nb_peer_comm *c = new nb_peer_comm();
  c->ctx = ctx; c->seq = 0; c->max_floats = max_floats;
  memset(&c->p, 0, sizeof(c->p));
  for (uint32_t r = 0; r < world; r++) {
    NB_REQUIRE(blocks[r], NB_ERR_ARG, "nb_peer_comm_create: block %u is NULL", r);
    c->p.base[r] = (uint8_t *)blocks[r];
  }
  c->p.rank = rank; c->p.world = world;
  c->p.slot_bytes = ((max_floats + 3) / 4 * 16 + 255) & ~255ull;
  NB_CUDA(cudaMalloc(&c->err_dev, 4));
  NB_CUDA(cudaMemsetAsync(c->err_dev, 0, 4, ctx->stream));
  // this rank's flags and slots start at zero; the caller barriers (host side) before the first exchange
  NB_CUDA(cudaMemsetAsync(blocks[rank], 0, nb_peer_comm_block_bytes(max_floats), ctx->stream));
  NB_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = c;
  return NB_OK;
}

int nb_peer_comm_destroy(nb_peer_comm *c) {
  if (!c) return NB_OK;
  DeviceGuard guard(c->ctx->device);
  cudaFree(c->err_dev);
  delete c;
  return NB_OK;
}

/* in place: inout[i] = sum over ranks of inout[i], on c's stream; every rank must call it with the same n, in the same order */
int nb_peer_allreduce_sum(nb_peer_comm *c, float *inout, uint64_t n) {
  NB_REQUIRE(c && inout, NB_ERR_ARG, "nb_peer_allreduce_sum: NULL argument");
  NB_REQUIRE(n > 0 && n <= c->max_floats, NB_ERR_ARG, "nb_peer_allreduce_sum: %llu floats exceed the communicator's %llu", (unsigned long long)n, (unsigned long long)c->max_floats);
  NB_REQUIRE(((uintptr_t)inout & 15) == 0, NB_ERR_ARG, "nb_peer_allreduce_sum: buffer must be 16-byte aligned");
  nb_ctx *ctx = c->ctx;
  NB_GUARD(ctx);
  c->seq++;
  k_peer_allreduce<<<PEER_CTAS, PEER_THREADS, 0, ctx->stream>>>(c->p, inout, (uint32_t)n, c->seq, c->err_dev);
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
}

/* 0 = every exchange so far completed; 1 = some peer never arrived within the kernel's time limit (synchronises the stream) */
int nb_peer_comm_check(nb_peer_comm *c, int *timed_out) {
  NB_REQUIRE(c && timed_out, NB_ERR_ARG, "nb_peer_comm_check: NULL argument");
  NB_GUARD(c->ctx);
  uint32_t e = 0;
  NB_CUDA(cudaMemcpyAsync(&e, c->err_dev, 4, cudaMemcpyDeviceToHost, c->ctx->stream));
  NB_CUDA(cudaStreamSynchronize(c->ctx->stream));
  *timed_out = (int)e;
  return NB_OK;
}

}  // extern "C"
