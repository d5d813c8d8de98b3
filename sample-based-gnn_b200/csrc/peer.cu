// peer.cu -- sum all-reduce of a small buffer over NVLink peer memory (our own kernels, no NCCL on the step's path).
//
// Replaces Parameter::reduce_multi_gpu_gradient -> NCCL_Communicator::AllReduce per tensor (core/NtsScheduler.hpp:830-836,
// cuda/ntsCUDAGraphOP.cu:173-200): the only exchange of the data-parallel path, ~330 KB of dense weight gradients per step --
// latency bound. Every rank owns one peer-shareable block (nb_vmm_alloc) mapped by all ranks:
//     [flags: world x PEER_CTAS u32][slot 0: world x slot_bytes][slot 1: world x slot_bytes]
// The exchange is PUSH based and split in two phases that a caller may separate in time:
//   begin (k_peer_exchange<PUSH>): block b writes chunk b of this rank's buffer into region `rank` of slot (seq & 1) of EVERY
//       rank's block (remote 16-byte stores over NVLink; the local copy is an ordinary store), __threadfence_system, then stores
//       seq into flag[rank][b] of every rank. Nothing waits: the kernel ends as soon as its stores are issued.
//   end   (k_peer_exchange<REDUCE>): block b waits until its LOCAL flag[p][b] of every rank p reached seq (bounded: a rank that
//       never arrives raises an error flag after ~20 s instead of hanging the GPU) and sums the `world` regions of the LOCAL slot
//       in rank order: the same order on every rank, so the result is bit-identical everywhere and run to run.
// All reads of the reduce phase are local HBM reads; whatever time passes between begin and end (the next batch's gather and
// bottom aggregation in the training loop, ~0.2 ms) absorbs rank skew. The push kernel never waits, so it may run on a side stream
// beside the caller's next kernels (default; 0.1805 -> 0.1743 ms per step at N=2, profiles/r2_exchange_ab_n2_side_push.txt); the reduce
// kernel, the only one that polls, runs in the caller's stream behind the local push: it never competes for SM slots with a
// persistent kernel of another stream (round 1's one-kernel rendezvous on a side stream could not start while the segment
// reduction owned every thread slot, and the other ranks spun for it: 0.52 scaling efficiency at N=8).
// nb_peer_allreduce_sum = both phases in one launch (k_peer_exchange<PUSH|REDUCE>).
// Two slots: rank r pushes exchange s+2 into slot (s & 1) only after its own reduce of s+1, which saw every peer's flag s+1,
// which every peer stores after its own reduce of s (stream order reduce(s) -> push(s+1)): nobody still reads slot (s & 1).
#include "common.cuh"

constexpr int PEER_CTAS = 32, PEER_THREADS = 256, PEER_MAX_WORLD = 16;
constexpr size_t PEER_FLAG_BYTES = 4096;   // world x PEER_CTAS u32, padded
constexpr int PEER_PUSH = 1, PEER_REDUCE = 2;

struct PeerParams {
  uint8_t *base[PEER_MAX_WORLD];   // every rank's block as mapped on THIS device
  uint32_t rank, world;
  uint64_t slot_bytes;             // one rank's region inside a slot
};

struct PeerStats {                 // device resident; wait = time the reduce phase spent polling for the slowest peer
  unsigned long long exchanges, wait_ns_sum, wait_ns_max;
  uint32_t err, pad;
};

struct nb_peer_comm {
  nb_ctx *ctx;
  PeerParams p;
  uint32_t seq;        // exchanges begun
  uint32_t seq_done;   // exchanges ended
  uint64_t max_floats;
  PeerStats *stats_dev;
  // push on a side stream ("peer_push_side_stream", default on): begin's kernel waits for nobody and only stores, so it can run
  // beside whatever the caller enqueues next instead of in front of it; end joins it before the reduce kernel
  cudaStream_t push_stream;
  cudaEvent_t ev_ready, ev_pushed;
};

static int g_peer_push_side = 1;
void nb_peer_set_push_side(int v) { g_peer_push_side = v ? 1 : 0; }

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <int PHASES>
__global__ void __launch_bounds__(PEER_THREADS)
k_peer_exchange(PeerParams c, const float *in, float *out, uint32_t n, uint32_t seq, PeerStats *stats) {   // in == out for the one-launch form
  __shared__ int s_timeout;
  __shared__ unsigned long long s_wait[PEER_MAX_WORLD];
  const uint32_t n4 = (n + 3) / 4;                                     // 16-byte units; regions are zero-padded to a multiple of 4 floats
  const uint32_t per = (n4 + gridDim.x - 1) / gridDim.x;
  const uint32_t lo = blockIdx.x * per, hi = min(n4, lo + per);
  const uint64_t slot_off = PEER_FLAG_BYTES + (uint64_t)(seq & 1u) * c.world * c.slot_bytes;
  if (PHASES & PEER_PUSH) {
    const uint64_t mine = slot_off + (uint64_t)c.rank * c.slot_bytes;
    for (uint32_t i = lo + threadIdx.x; i < hi; i += PEER_THREADS) {
      float4 v;
      if (4 * i + 3 < n) v = *reinterpret_cast<const float4 *>(in + 4 * (uint64_t)i);
      else {
        v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (4 * i < n) v.x = in[4 * (uint64_t)i];
        if (4 * i + 1 < n) v.y = in[4 * (uint64_t)i + 1];
        if (4 * i + 2 < n) v.z = in[4 * (uint64_t)i + 2];
      }
      for (uint32_t q = 0; q < c.world; q++) {
        const uint32_t p = (c.rank + q) % c.world;                     // start with the local copy, spread the peers over the links
        *reinterpret_cast<float4 *>(c.base[p] + mine + 16 * (uint64_t)i) = v;
      }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < c.world)   // chunk b of this rank is in place on rank p
      *(reinterpret_cast<volatile uint32_t *>(c.base[threadIdx.x]) + c.rank * PEER_CTAS + blockIdx.x) = seq;
  }
  if (PHASES & PEER_REDUCE) {
    if (threadIdx.x == 0) s_timeout = 0;
    __syncthreads();
    if (threadIdx.x < c.world) {
      volatile uint32_t *flag = reinterpret_cast<volatile uint32_t *>(c.base[c.rank]) + threadIdx.x * PEER_CTAS + blockIdx.x;
      const unsigned long long t0 = globaltimer_ns();
      unsigned long long waited = 0;
      while ((int)(*flag - seq) < 0) {
        waited = globaltimer_ns() - t0;
        if (waited > 20000000000ull) { s_timeout = 1; break; }         // 20 s: a peer never arrived
      }
      s_wait[threadIdx.x] = waited;
      __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long w = 0;
      for (uint32_t p = 0; p < c.world; p++) w = max(w, s_wait[p]);
      atomicMax(&stats->wait_ns_max, w);
      if (blockIdx.x == 0) { atomicAdd(&stats->wait_ns_sum, w); atomicAdd(&stats->exchanges, 1ull); }
      if (s_timeout) atomicExch(&stats->err, 1u);
    }
    if (s_timeout) return;
    const float *slot = reinterpret_cast<const float *>(c.base[c.rank] + slot_off);
    const uint64_t region = c.slot_bytes / 4;
    for (uint32_t i = lo + threadIdx.x; i < hi; i += PEER_THREADS) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      // the regions were written by the peers before the flags this block has just observed (volatile poll, system fence, barrier):
      // L1-bypassing loads see them, and unlike volatile ones eight of them are in flight at once (volatile loads issue one after
      // the other: 8 ranks x 3 iterations of L2 latency made the reduce phase 13 us of the step at N=8)
      for (uint32_t p0 = 0; p0 < c.world; p0 += 8) {
        float4 v[8];
#pragma unroll
        for (int k = 0; k < 8; k++)
          if (p0 + k < c.world) v[k] = __ldcg(reinterpret_cast<const float4 *>(slot + (p0 + k) * region) + i);
#pragma unroll
        for (int k = 0; k < 8; k++)
          if (p0 + k < c.world) { acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w; }   // rank order
      }
      if (4 * i + 3 < n) *reinterpret_cast<float4 *>(out + 4 * (uint64_t)i) = acc;
      else {
        if (4 * i < n) out[4 * (uint64_t)i] = acc.x;
        if (4 * i + 1 < n) out[4 * (uint64_t)i + 1] = acc.y;
        if (4 * i + 2 < n) out[4 * (uint64_t)i + 2] = acc.z;
      }
    }
  }
}

static int peer_launch(nb_peer_comm *c, int phases, const float *in, float *out, uint64_t n) {
  nb_ctx *ctx = c->ctx;
  const uint32_t seq = (phases & PEER_PUSH) ? c->seq : c->seq_done;
  if (c->push_stream && phases == PEER_PUSH) {          // fork: the push runs beside the caller's next kernels
    NB_CUDA(cudaEventRecord(c->ev_ready, ctx->stream));
    NB_CUDA(cudaStreamWaitEvent(c->push_stream, c->ev_ready, 0));
    k_peer_exchange<PEER_PUSH><<<PEER_CTAS, PEER_THREADS, 0, c->push_stream>>>(c->p, in, out, (uint32_t)n, seq, c->stats_dev);
    NB_LAUNCH_CHECK(ctx);
    NB_CUDA(cudaEventRecord(c->ev_pushed, c->push_stream));
    return NB_OK;
  }
  if (c->push_stream && phases == PEER_REDUCE) NB_CUDA(cudaStreamWaitEvent(ctx->stream, c->ev_pushed, 0));   // join
  if (phases == (PEER_PUSH | PEER_REDUCE)) k_peer_exchange<PEER_PUSH | PEER_REDUCE><<<PEER_CTAS, PEER_THREADS, 0, ctx->stream>>>(c->p, in, out, (uint32_t)n, seq, c->stats_dev);
  else if (phases == PEER_PUSH) k_peer_exchange<PEER_PUSH><<<PEER_CTAS, PEER_THREADS, 0, ctx->stream>>>(c->p, in, out, (uint32_t)n, seq, c->stats_dev);
  else k_peer_exchange<PEER_REDUCE><<<PEER_CTAS, PEER_THREADS, 0, ctx->stream>>>(c->p, in, out, (uint32_t)n, seq, c->stats_dev);
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
}

extern "C" {

size_t nb_peer_comm_block_bytes(uint64_t max_floats, uint32_t world) {
  const uint64_t region = ((max_floats + 3) / 4 * 16 + 255) & ~255ull;
  return PEER_FLAG_BYTES + 2 * (size_t)(world ? world : 1) * region;
}

int nb_peer_comm_create(nb_ctx *ctx, uint32_t rank, uint32_t world, uint64_t max_floats, void *const *blocks, nb_peer_comm **out) {
  NB_REQUIRE(ctx && out && blocks && world >= 1 && world <= PEER_MAX_WORLD && rank < world && max_floats > 0, NB_ERR_ARG, "nb_peer_comm_create: bad argument");
  NB_REQUIRE((size_t)world * PEER_CTAS * 4 <= PEER_FLAG_BYTES, NB_ERR_ARG, "nb_peer_comm_create: world too large");
  NB_GUARD(ctx);
  nb_peer_comm *c = new nb_peer_comm();
  c->ctx = ctx; c->seq = 0; c->seq_done = 0; c->max_floats = max_floats;
  memset(&c->p, 0, sizeof(c->p));
  for (uint32_t r = 0; r < world; r++) {
    NB_REQUIRE(blocks[r], NB_ERR_ARG, "nb_peer_comm_create: block %u is NULL", r);
    c->p.base[r] = (uint8_t *)blocks[r];
  }
  c->p.rank = rank; c->p.world = world;
  c->p.slot_bytes = ((max_floats + 3) / 4 * 16 + 255) & ~255ull;
  c->push_stream = nullptr; c->ev_ready = nullptr; c->ev_pushed = nullptr;
  if (g_peer_push_side && world > 1) {
    int lo = 0, hi = 0;
    NB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    NB_CUDA(cudaStreamCreateWithPriority(&c->push_stream, cudaStreamNonBlocking, hi));
    NB_CUDA(cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming));
    NB_CUDA(cudaEventCreateWithFlags(&c->ev_pushed, cudaEventDisableTiming));
  }
  NB_CUDA(cudaMalloc(&c->stats_dev, sizeof(PeerStats)));
  NB_CUDA(cudaMemsetAsync(c->stats_dev, 0, sizeof(PeerStats), ctx->stream));
  // this rank's flags start at zero; the caller barriers (host side) before the first exchange
  NB_CUDA(cudaMemsetAsync(blocks[rank], 0, PEER_FLAG_BYTES, ctx->stream));
  NB_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = c;
  return NB_OK;
}

int nb_peer_comm_destroy(nb_peer_comm *c) {
  if (!c) return NB_OK;
  DeviceGuard guard(c->ctx->device);
  if (c->push_stream) { cudaStreamSynchronize(c->push_stream); cudaStreamDestroy(c->push_stream); cudaEventDestroy(c->ev_ready); cudaEventDestroy(c->ev_pushed); }
  cudaFree(c->stats_dev);
  delete c;
  return NB_OK;
}

static int peer_args(nb_peer_comm *c, const float *buf, uint64_t n, const char *who) {
  NB_REQUIRE(c && buf, NB_ERR_ARG, "%s: NULL argument", who);
  NB_REQUIRE(n > 0 && n <= c->max_floats, NB_ERR_ARG, "%s: %llu floats exceed the communicator's %llu", who, (unsigned long long)n, (unsigned long long)c->max_floats);
  NB_REQUIRE(((uintptr_t)buf & 15) == 0, NB_ERR_ARG, "%s: buffer must be 16-byte aligned", who);
  return NB_OK;
}

/* in place: inout[i] = sum over ranks of inout[i], on c's stream; every rank must call it with the same n, in the same order */
int nb_peer_allreduce_sum(nb_peer_comm *c, float *inout, uint64_t n) {
  int rc = peer_args(c, inout, n, "nb_peer_allreduce_sum");
  if (rc) return rc;
  NB_REQUIRE(c->seq == c->seq_done, NB_ERR_ARG, "nb_peer_allreduce_sum: an exchange begun with nb_peer_allreduce_begin is still open");
  NB_GUARD(c->ctx);
  c->seq++; c->seq_done++;
  return peer_launch(c, PEER_PUSH | PEER_REDUCE, inout, inout, n);
}

int nb_peer_allreduce_begin(nb_peer_comm *c, const float *in, uint64_t n) {
  int rc = peer_args(c, in, n, "nb_peer_allreduce_begin");
  if (rc) return rc;
  NB_REQUIRE(c->seq == c->seq_done, NB_ERR_ARG, "nb_peer_allreduce_begin: the previous exchange has not been ended (one exchange in flight per communicator)");
  NB_GUARD(c->ctx);
  c->seq++;
  return peer_launch(c, PEER_PUSH, in, nullptr, n);
}

int nb_peer_allreduce_end(nb_peer_comm *c, float *out, uint64_t n) {
  int rc = peer_args(c, out, n, "nb_peer_allreduce_end");
  if (rc) return rc;
  NB_REQUIRE(c->seq == c->seq_done + 1, NB_ERR_ARG, "nb_peer_allreduce_end: no exchange in flight");
  NB_GUARD(c->ctx);
  c->seq_done++;
  return peer_launch(c, PEER_REDUCE, nullptr, out, n);
}

/* 0 = every exchange so far completed; 1 = some peer never arrived within the kernel's time limit (synchronises the stream) */
int nb_peer_comm_check(nb_peer_comm *c, int *timed_out) {
  NB_REQUIRE(c && timed_out, NB_ERR_ARG, "nb_peer_comm_check: NULL argument");
  NB_GUARD(c->ctx);
  PeerStats h;
  NB_CUDA(cudaMemcpyAsync(&h, c->stats_dev, sizeof(h), cudaMemcpyDeviceToHost, c->ctx->stream));
  NB_CUDA(cudaStreamSynchronize(c->ctx->stream));
  *timed_out = (int)h.err;
  return NB_OK;
}

/* time the reduce phases spent polling for the slowest peer since the last reset (synchronises the stream) */
int nb_peer_comm_stats(nb_peer_comm *c, uint64_t *exchanges, uint64_t *wait_ns_sum, uint64_t *wait_ns_max, int reset) {
  NB_REQUIRE(c, NB_ERR_ARG, "nb_peer_comm_stats: NULL argument");
  NB_GUARD(c->ctx);
  PeerStats h;
  NB_CUDA(cudaMemcpyAsync(&h, c->stats_dev, sizeof(h), cudaMemcpyDeviceToHost, c->ctx->stream));
  if (reset) NB_CUDA(cudaMemsetAsync(c->stats_dev, 0, offsetof(PeerStats, err), c->ctx->stream));
  NB_CUDA(cudaStreamSynchronize(c->ctx->stream));
  if (exchanges) *exchanges = h.exchanges;
  if (wait_ns_sum) *wait_ns_sum = h.wait_ns_sum;
  if (wait_ns_max) *wait_ns_max = h.wait_ns_max;
  return NB_OK;
}

}  // extern "C"
