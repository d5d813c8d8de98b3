// stage.cu -- cold-row staging for feature tables that stay in host memory.
//
// Hot rows live in an HBM cache table (cache_node_hashmap[v] = slot or -1, top-out-degree vertices:
// toolkits/GS_SAMPLE_PC_MULTI.hpp:916-1015); cold rows come from the host table. The reference splits the id list on the CPU
// after a synchronous D2H and then reads cold rows over PCIe 4 bytes at a time from a zero-copy mapping
// (core/ntsFastSampler.hpp:263-317). Here the split happens on the device, only the cold ids travel to the host, a worker
// thread packs those rows into a pinned staging buffer and ships them with one cudaMemcpyAsync on a side stream, and the merge
// (the gather kernel in its three-tier mode, gather.cu: hot rows from the cache, cold rows from the staged block) waits on that
// copy's event. submit() returns immediately, so the staging of batch i+1 overlaps the training of batch i (2 slots).
// The cache table may itself be sharded over the GPUs of the node and read over NVLink (nb_stage_gather_table): that is the
// papers100M-shaped configuration -- partitioned hot cache + host-streamed cold rows.
#include <condition_variable>
#include <mutex>
#include <thread>
#include <atomic>
#include <vector>
#include <stdlib.h>

#include "common.cuh"
#include "pack_pool.h"

constexpr int STAGE_SLOTS = 2;

struct StageSlot {
  uint32_t *cold_slot_dev, *cold_ids_dev, *count_dev;  // cold rows: batch position -> row of the staged block; compacted global ids
  uint32_t *cold_ids_host, *count_host;               // pinned
  float *rows_host, *rows_dev;                        // pinned staging block and its device copy
  cudaEvent_t ids_ready, rows_ready, consumed;
  uint32_t n_rows, n_cold;
  int state;  // 0 idle, 1 submitted (worker owns it), 2 staged (rows_ready recorded)
};

struct PackPool;
struct nb_stage {
  nb_ctx *ctx;
  cudaStream_t side;
  const float *host_table;
  uint64_t host_pitch;
  uint32_t F, max_rows;
  StageSlot slot[STAGE_SLOTS];
  std::thread worker;
  PackPool *pool;
  std::mutex m;
  std::condition_variable cv;
  int pending[STAGE_SLOTS + 1], n_pending;
  bool stop;
  char err[256];
};

// rows whose cache slot is -1: warp-ballot compaction, one atomic per warp. The order of the list does not matter to the
// merge (each batch position remembers which staged row is its own), so no scan is needed
__global__ void __launch_bounds__(256)
k_cold_split(const uint32_t *__restrict__ ids, const uint32_t *__restrict__ cache_map, uint32_t n, uint32_t *__restrict__ cold_slot,
             uint32_t *__restrict__ cold_ids, uint32_t *count) {
  const unsigned lane = lane_id();
  for (unsigned i0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; i0 < n; i0 += gridDim.x * blockDim.x) {
    const unsigned i = i0 + lane;
    uint32_t v = 0;
    bool cold = false;
    if (i < n) { v = ids[i]; cold = cache_map[v] == 0xffffffffu; }
    const unsigned mask = __ballot_sync(FULL_MASK, cold);
    if (!mask) continue;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(count, (uint32_t)__popc(mask));
    base = __shfl_sync(FULL_MASK, base, 0);
    if (cold) {
      const uint32_t k = base + __popc(mask & ((1u << lane) - 1u));
      cold_slot[i] = k;
      cold_ids[k] = v;
    }
  }
}

// Threads packing cold rows: NB_STAGE_THREADS, else hardware threads / visible GPUs clamped to [2, 16] (one process per GPU:
// the ranks of a node share the host's cores). A private pool that sleeps on a condition variable between batches: an OpenMP
// team here spins while idle, and with one team per rank on a shared host that oversubscribes the cores (measured: 4 ranks x 16
// OpenMP threads on 32 cores packed 2.5 GB/s per rank against 22 GB/s for a single rank).
static int stage_threads() {
  static int n = 0;
  if (!n) {
    const char *e = getenv("NB_STAGE_THREADS");
    if (e) n = atoi(e);
    else {
      int gpus = 1;
      if (cudaGetDeviceCount(&gpus) != cudaSuccess || gpus < 1) { cudaGetLastError(); gpus = 1; }
      n = (int)std::thread::hardware_concurrency() / gpus;
      if (n > 16) n = 16;
      if (n < 2) n = 2;
    }
    if (n < 1) n = 1;
  }
  return n;
}

static void stage_worker(nb_stage *s) {
  cudaSetDevice(s->ctx->device);
  while (true) {
    int k;
    {
      std::unique_lock<std::mutex> g(s->m);
      s->cv.wait(g, [&] { return s->stop || s->n_pending > 0; });
      if (s->stop && s->n_pending == 0) return;
      k = s->pending[0];
      for (int i = 1; i < s->n_pending; i++) s->pending[i - 1] = s->pending[i];
      s->n_pending--;
    }
    StageSlot &sl = s->slot[k];
    const bool tr = nb_trace_on();
    uint64_t t0 = tr ? nb_trace_now_ns() : 0;
    cudaError_t e = cudaEventSynchronize(sl.ids_ready);  // cold ids and their count are on the host
    if (tr) { const uint64_t t1 = nb_trace_now_ns(); nb_trace_add("stage_worker: wait for cold ids", t1 - t0); t0 = t1; }
    const uint32_t nc = e == cudaSuccess ? *sl.count_host : 0;
    if (e == cudaSuccess) {
      const size_t row_bytes = (size_t)s->F * sizeof(float);
      if (nc) s->pool->pack(s->host_table, s->host_pitch, sl.cold_ids_host, sl.rows_host, nc, s->F);
      if (tr) { const uint64_t t1 = nb_trace_now_ns(); nb_trace_add("stage_worker: pack cold rows", t1 - t0); t0 = t1; }
      if (nc) e = cudaMemcpyAsync(sl.rows_dev, sl.rows_host, (size_t)nc * row_bytes, cudaMemcpyHostToDevice, s->side);
      if (e == cudaSuccess) e = cudaEventRecord(sl.rows_ready, s->side);
      if (tr) {
        cudaEventSynchronize(sl.rows_ready);   // tracing only: makes the copy's duration visible (and serialises it)
        nb_trace_add("stage_worker: H2D copy of the packed rows", nb_trace_now_ns() - t0);
      }
    }
    {
      std::lock_guard<std::mutex> g(s->m);
      sl.n_cold = nc;
      sl.state = 2;
      if (e != cudaSuccess) snprintf(s->err, sizeof(s->err), "stage worker: %s", cudaGetErrorString(e));
    }
    s->cv.notify_all();
  }
}

extern "C" {

int nb_stage_create(nb_ctx *ctx, const float *host_table, uint32_t host_pitch, uint32_t feature_size, uint32_t max_rows, nb_stage **out) {
  NB_REQUIRE(ctx && host_table && out && feature_size > 0 && host_pitch >= feature_size && max_rows > 0, NB_ERR_ARG, "nb_stage_create: bad argument");
  NB_GUARD(ctx);
  nb_stage *s = new nb_stage();
  s->ctx = ctx; s->host_table = host_table; s->host_pitch = host_pitch; s->F = feature_size; s->max_rows = max_rows;
  s->n_pending = 0; s->stop = false; s->err[0] = 0;
  NB_CUDA(cudaStreamCreateWithFlags(&s->side, cudaStreamNonBlocking));
  for (int k = 0; k < STAGE_SLOTS; k++) {
    StageSlot &sl = s->slot[k];
    NB_CUDA(cudaMalloc(&sl.cold_slot_dev, (size_t)max_rows * 4));
    NB_CUDA(cudaMalloc(&sl.cold_ids_dev, (size_t)max_rows * 4));
    NB_CUDA(cudaMalloc(&sl.count_dev, 64));
    NB_CUDA(cudaMalloc(&sl.rows_dev, (size_t)max_rows * feature_size * 4));
    NB_CUDA(cudaHostAlloc(&sl.cold_ids_host, (size_t)max_rows * 4, cudaHostAllocDefault));
    NB_CUDA(cudaHostAlloc(&sl.count_host, 64, cudaHostAllocDefault));
    NB_CUDA(cudaHostAlloc(&sl.rows_host, (size_t)max_rows * feature_size * 4, cudaHostAllocDefault));
    NB_CUDA(cudaEventCreateWithFlags(&sl.ids_ready, cudaEventDisableTiming));
    NB_CUDA(cudaEventCreateWithFlags(&sl.rows_ready, cudaEventDisableTiming));
    NB_CUDA(cudaEventCreateWithFlags(&sl.consumed, cudaEventDisableTiming));
    sl.state = 0; sl.n_rows = 0; sl.n_cold = 0;
  }
  s->pool = new PackPool();
  s->pool->start(stage_threads());
  s->worker = std::thread(stage_worker, s);
  *out = s;
  return NB_OK;
}

int nb_stage_destroy(nb_stage *s) {
  if (!s) return NB_OK;
  { std::lock_guard<std::mutex> g(s->m); s->stop = true; }
  s->cv.notify_all();
  if (s->worker.joinable()) s->worker.join();
  s->pool->shutdown();
  delete s->pool;
  DeviceGuard guard(s->ctx->device);
  cudaStreamSynchronize(s->side);
  for (int k = 0; k < STAGE_SLOTS; k++) {
    StageSlot &sl = s->slot[k];
    cudaFree(sl.cold_slot_dev); cudaFree(sl.cold_ids_dev); cudaFree(sl.count_dev); cudaFree(sl.rows_dev);
    cudaFreeHost(sl.cold_ids_host); cudaFreeHost(sl.count_host); cudaFreeHost(sl.rows_host);
    cudaEventDestroy(sl.ids_ready); cudaEventDestroy(sl.rows_ready); cudaEventDestroy(sl.consumed);
  }
  cudaStreamDestroy(s->side);
  delete s;
  return NB_OK;
}

// Enqueue the split of one id list on the ctx stream and hand the slot to the worker. Returns immediately.
int nb_stage_submit(nb_stage *s, int slot, const uint32_t *ids_dev, uint32_t n_rows, const uint32_t *cache_node_hashmap_dev) {
  NB_REQUIRE(s && slot >= 0 && slot < STAGE_SLOTS && (n_rows == 0 || (ids_dev && cache_node_hashmap_dev)), NB_ERR_ARG, "nb_stage_submit: bad argument");
  NB_REQUIRE(n_rows <= s->max_rows, NB_ERR_CAPACITY, "nb_stage_submit: %u rows exceed the stage capacity %u", n_rows, s->max_rows);
  nb_ctx *ctx = s->ctx;
  NB_GUARD(ctx);
  StageSlot &sl = s->slot[slot];
  {
    std::unique_lock<std::mutex> g(s->m);
    NB_REQUIRE(sl.state != 1, NB_ERR_ARG, "nb_stage_submit: slot %d is still being staged", slot);
    sl.state = 1;
  }
  // any failure below must leave the slot free again: a slot stuck in state 1 that was never queued would make the next
  // nb_stage_gather wait for ever
  auto enqueue = [&]() -> int {
    NB_CUDA(cudaStreamWaitEvent(ctx->stream, sl.consumed, 0));  // the merge that last read this slot's buffers has run
    NB_CUDA(cudaMemsetAsync(sl.count_dev, 0, 4, ctx->stream));
    if (n_rows) {
      k_cold_split<<<nb_grid(n_rows, 256, 4), 256, 0, ctx->stream>>>(ids_dev, cache_node_hashmap_dev, n_rows, sl.cold_slot_dev, sl.cold_ids_dev, sl.count_dev);
      NB_LAUNCH_CHECK(ctx);
      NB_CUDA(cudaMemcpyAsync(sl.cold_ids_host, sl.cold_ids_dev, (size_t)n_rows * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    NB_CUDA(cudaMemcpyAsync(sl.count_host, sl.count_dev, 4, cudaMemcpyDeviceToHost, ctx->stream));
    NB_CUDA(cudaEventRecord(sl.ids_ready, ctx->stream));
    return NB_OK;
  };
  const int rc = enqueue();
  if (rc != NB_OK) {
    std::lock_guard<std::mutex> g(s->m);
    sl.state = 0;
    return rc;
  }
  sl.n_rows = n_rows;
  {
    std::lock_guard<std::mutex> g(s->m);
    s->pending[s->n_pending++] = slot;
  }
  s->cv.notify_all();
  return NB_OK;
}

// out[i,:] for the id list given to submit(): waits (host side) until the worker has issued the copy, then orders the merge
// kernel behind it on the ctx stream. n_cold_out (may be NULL) receives the number of rows that came from the host.
static int stage_gather(nb_stage *s, int slot, float *out, uint32_t out_pitch, const float *cache_table, uint32_t cache_pitch,
                        const nb_table *hot, const uint32_t *cache_node_hashmap_dev, const uint32_t *ids_dev, uint32_t *n_cold_out) {
  nb_ctx *ctx = s->ctx;
  NB_GUARD(ctx);
  StageSlot &sl = s->slot[slot];
  {
    std::unique_lock<std::mutex> g(s->m);
    NB_REQUIRE(sl.state != 0, NB_ERR_ARG, "nb_stage_gather: nothing was submitted on slot %d", slot);
    s->cv.wait(g, [&] { return sl.state == 2; });
    sl.state = 0;
    if (s->err[0]) { nb_set_error("%s", s->err); s->err[0] = 0; return NB_ERR_CUDA; }
  }
  if (n_cold_out) *n_cold_out = sl.n_cold;
  NB_CUDA(cudaStreamWaitEvent(ctx->stream, sl.rows_ready, 0));
  if (sl.n_rows) {
    int rc = nb_launch_gather_tiered(ctx, out, out_pitch, cache_table, cache_pitch, hot, cache_node_hashmap_dev, ids_dev, sl.n_rows,
                                     sl.rows_dev, s->F, sl.cold_slot_dev, s->F);
    if (rc != NB_OK) return rc;
  }
  NB_CUDA(cudaEventRecord(sl.consumed, ctx->stream));
  return NB_OK;
}

int nb_stage_gather(nb_stage *s, int slot, float *out, uint32_t out_pitch, const float *cache_table, uint32_t cache_pitch,
                    const uint32_t *cache_node_hashmap_dev, const uint32_t *ids_dev, uint32_t *n_cold_out) {
  NB_REQUIRE(s && slot >= 0 && slot < STAGE_SLOTS && out && out_pitch >= s->F, NB_ERR_ARG, "nb_stage_gather: bad argument");
  NB_REQUIRE(cache_pitch >= s->F, NB_ERR_ARG, "nb_stage_gather: bad cache pitch");
  return stage_gather(s, slot, out, out_pitch, cache_table, cache_pitch, nullptr, cache_node_hashmap_dev, ids_dev, n_cold_out);
}

int nb_stage_gather_table(nb_stage *s, int slot, float *out, uint32_t out_pitch, nb_table *hot_table,
                          const uint32_t *cache_node_hashmap_dev, const uint32_t *ids_dev, uint32_t *n_cold_out) {
  NB_REQUIRE(s && slot >= 0 && slot < STAGE_SLOTS && out && out_pitch >= s->F && hot_table, NB_ERR_ARG, "nb_stage_gather_table: bad argument");
  NB_REQUIRE(hot_table->feature_size == s->F, NB_ERR_ARG, "nb_stage_gather_table: the hot table holds rows of %u floats, the stage of %u",
             hot_table->feature_size, s->F);
  return stage_gather(s, slot, out, out_pitch, nullptr, 0, hot_table, cache_node_hashmap_dev, ids_dev, n_cold_out);
}

}  // extern "C"
