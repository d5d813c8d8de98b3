// ctx.cu -- context, error reporting and memory helpers of the C ABI (include/nts_b200.h).
// Replaces class Cuda_Stream's stream ownership (cuda/ntsCUDAGraphOP.cu:203-262 of the reference)
// and the free allocation/copy helpers (cuda/ntsCUDA.hpp:30-71).
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void nb_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int nb_ctx_scratch(nb_ctx *ctx, size_t bytes, void **out) {
  if (bytes > ctx->scratch_bytes) {
    // kernels still using the old buffer are ordered before this point on the stream
    NB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->scratch) NB_CUDA(cudaFree(ctx->scratch));
    ctx->scratch = nullptr;
    ctx->scratch_bytes = 0;
    size_t want = bytes + bytes / 4 + 4096;
    NB_CUDA(cudaMalloc(&ctx->scratch, want));
    ctx->scratch_bytes = want;
  }
  *out = ctx->scratch;
  return NB_OK;
}

#include <stdlib.h>
// ---- HBM mirror of host-resident tables (features, adjacency) --------------------------------------
// The reference keeps the feature table and the adjacency (row_indices) in mapped pinned host memory and reads them over
// PCIe (zero copy: core/ntsDataloador.hpp:187,483; core/FullyRepGraph.hpp:727 + core/ntsFastSampler.hpp:159-166). On a 180 GB part the table fits in HBM, so the first
// gather that sees a host-resident table copies the whole allocation to the device once and every later gather reads
// HBM. The table is treated as immutable after that first gather (it is, in every sampled toolkit).
// Opt-in for feature tables (NB_MIRROR_HOST_TABLES=1 / nb_set_option("mirror_host_tables", 1)): some toolkits also push
// host buffers that the CPU rewrites every super-batch through the same call. On by default for the adjacency.
#include <map>
#include <mutex>
static int g_mirror_tables = -1;
static std::mutex g_mirror_mutex;
struct MirrorEntry { uintptr_t dev_base; size_t size; void *mirror; };
static std::map<std::pair<int, uintptr_t>, MirrorEntry> g_mirrors;  // (device, device-visible base of the host allocation)

static int g_mirror_adjacency = -1;
const void *nb_mirror_host(nb_ctx *ctx, const void *table, int is_adjacency) {
  if (g_mirror_tables < 0) {
    const char *e = getenv("NB_MIRROR_HOST_TABLES");
    g_mirror_tables = e ? atoi(e) : 0;   // feature-like buffers may be rewritten by the host (CPU-computed hot embeddings): opt-in
  }
  if (g_mirror_adjacency < 0) {
    const char *e = getenv("NB_MIRROR_HOST_ADJACENCY");
    g_mirror_adjacency = e ? atoi(e) : 1;  // the topology never changes after load: on by default
  }
  if (!(is_adjacency ? g_mirror_adjacency : g_mirror_tables) || !table) return table;
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, table) != cudaSuccess) { cudaGetLastError(); return table; }
  if (attr.type != cudaMemoryTypeHost || !attr.devicePointer) return table;
  typedef int (*range_fn)(unsigned long long *, size_t *, unsigned long long);
  static range_fn get_range = nullptr;
  if (!get_range) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) { cudaGetLastError(); return table; }
    get_range = (range_fn)fn;
  }
  unsigned long long base = 0;
  size_t size = 0;
  if (get_range(&base, &size, (unsigned long long)(uintptr_t)attr.devicePointer) != 0 || !size) return table;
  std::lock_guard<std::mutex> lock(g_mirror_mutex);
  auto key = std::make_pair(ctx->device, (uintptr_t)base);
  auto it = g_mirrors.find(key);
  if (it == g_mirrors.end()) {
    void *m = nullptr;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || size > free_b / 2 || cudaMalloc(&m, size) != cudaSuccess) {
      cudaGetLastError();
      g_mirrors[key] = MirrorEntry{(uintptr_t)base, size, nullptr};  // remember the refusal, stay zero-copy
      return table;
    }
    if (cudaMemcpyAsync(m, (const void *)(uintptr_t)base, size, cudaMemcpyDefault, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) { cudaGetLastError(); cudaFree(m); return table; }
    it = g_mirrors.insert(std::make_pair(key, MirrorEntry{(uintptr_t)base, size, m})).first;
  }
  if (!it->second.mirror) return table;
  return (const void *)((const char *)it->second.mirror + ((uintptr_t)attr.devicePointer - it->second.dev_base));
}
void nb_mirror_host_enable(int on, int adjacency) { if (adjacency) g_mirror_adjacency = on; else g_mirror_tables = on; }


extern "C" {

int nb_abi_version(void) { return NB_ABI_VERSION; }
const char *nb_last_error(void) { return g_err; }

int nb_device_count(int *count) {
  NB_REQUIRE(count, NB_ERR_ARG, "count is NULL");
  *count = 0;
  NB_CUDA(cudaGetDeviceCount(count));
  return NB_OK;
}

int nb_ctx_create(int device, void *cuda_stream, int adopt_stream, nb_ctx **out) {
  NB_REQUIRE(out, NB_ERR_ARG, "out is NULL");
  *out = nullptr;
  int n = 0;
  NB_CUDA(cudaGetDeviceCount(&n));
  NB_REQUIRE(device >= 0 && device < n, NB_ERR_ARG, "device %d out of range (%d devices)", device, n);
  nb_ctx *c = new nb_ctx();
  c->device = device;
  c->launches = 0;
  c->scratch = nullptr;
  c->scratch_bytes = 0;
  DeviceGuard g(device);
  if (!g.ok) { delete c; nb_set_error("cudaSetDevice(%d) failed", device); return NB_ERR_CUDA; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; nb_set_error("cudaGetDeviceProperties failed"); return NB_ERR_CUDA; }
  if (prop.major != 10) {
    delete c;
    nb_set_error("device %d is sm_%d%d; libnts_b200 is built for sm_100a only", device, prop.major, prop.minor);
    return NB_ERR_UNSUPPORTED;
  }
  c->sm_count = prop.multiProcessorCount;
  if (adopt_stream) {
    c->stream = (cudaStream_t)cuda_stream;
    c->own_stream = false;
  } else {
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
      delete c; nb_set_error("cudaStreamCreate failed"); return NB_ERR_CUDA;
    }
    c->own_stream = true;
  }
  *out = c;
  return NB_OK;
}

int nb_ctx_destroy(nb_ctx *ctx) {
  if (!ctx) return NB_OK;
  DeviceGuard g(ctx->device);
  if (ctx->scratch) cudaFree(ctx->scratch);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return NB_OK;
}

int nb_ctx_set_stream(nb_ctx *ctx, void *cuda_stream) {
  NB_REQUIRE(ctx, NB_ERR_ARG, "ctx is NULL");
  NB_GUARD(ctx);
  if (ctx->own_stream) { NB_CUDA(cudaStreamSynchronize(ctx->stream)); NB_CUDA(cudaStreamDestroy(ctx->stream)); }
  ctx->stream = (cudaStream_t)cuda_stream;
  ctx->own_stream = false;
  return NB_OK;
}

void *nb_ctx_stream(nb_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
int nb_ctx_device(nb_ctx *ctx) { return ctx ? ctx->device : -1; }
uint64_t nb_ctx_launch_count(nb_ctx *ctx) { return ctx ? ctx->launches : 0; }

int nb_ctx_synchronize(nb_ctx *ctx) {
  NB_REQUIRE(ctx, NB_ERR_ARG, "ctx is NULL");
  NB_GUARD(ctx);
  NB_CUDA(cudaStreamSynchronize(ctx->stream));
  return NB_OK;
}

int nb_malloc_pinned(size_t bytes, void **out) {
  NB_REQUIRE(out, NB_ERR_ARG, "out is NULL");
  NB_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocMapped | cudaHostAllocPortable));
  return NB_OK;
}
int nb_free_host(void *p) { if (p) NB_CUDA(cudaFreeHost(p)); return NB_OK; }
int nb_device_pointer(void *host_mapped, void **out) {
  NB_REQUIRE(out, NB_ERR_ARG, "out is NULL");
  NB_CUDA(cudaHostGetDevicePointer(out, host_mapped, 0));
  return NB_OK;
}
int nb_malloc_device(size_t bytes, void **out) {
  NB_REQUIRE(out, NB_ERR_ARG, "out is NULL");
  NB_CUDA(cudaMalloc(out, bytes ? bytes : 1));
  return NB_OK;
}
int nb_free_device(void *p) { if (p) NB_CUDA(cudaFree(p)); return NB_OK; }

int nb_memcpy_h2d(nb_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes, int sync) {
  NB_REQUIRE(ctx, NB_ERR_ARG, "ctx is NULL");
  NB_GUARD(ctx);
  NB_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  if (sync) NB_CUDA(cudaStreamSynchronize(ctx->stream));
  return NB_OK;
}
int nb_memcpy_d2h(nb_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes, int sync) {
  NB_REQUIRE(ctx, NB_ERR_ARG, "ctx is NULL");
  NB_GUARD(ctx);
  NB_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  if (sync) NB_CUDA(cudaStreamSynchronize(ctx->stream));
  return NB_OK;
}
int nb_memset_async(nb_ctx *ctx, void *dst_dev, int value, size_t bytes) {
  NB_REQUIRE(ctx, NB_ERR_ARG, "ctx is NULL");
  NB_GUARD(ctx);
  NB_CUDA(cudaMemsetAsync(dst_dev, value, bytes, ctx->stream));
  return NB_OK;
}

int nb_ipc_get_handle(void *dev_ptr, void *handle64_out) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  NB_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle64_out, dev_ptr));
  return NB_OK;
}
int nb_ipc_open_handle(const void *handle64, void **dev_ptr_out) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  NB_CUDA(cudaIpcOpenMemHandle(dev_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return NB_OK;
}
int nb_ipc_close_handle(void *dev_ptr) { NB_CUDA(cudaIpcCloseMemHandle(dev_ptr)); return NB_OK; }

}  // extern "C"
