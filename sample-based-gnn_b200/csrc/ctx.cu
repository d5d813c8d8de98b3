// ctx.cu -- context, error reporting and memory helpers of the C ABI (include/nts_b200.h).
// Replaces class Cuda_Stream's stream ownership (cuda/ntsCUDAGraphOP.cu:203-262 of the reference)
// and the free allocation/copy helpers (cuda/ntsCUDA.hpp:30-71).
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void nb_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- tracing ------------------------------------------------------------------------------------------
#include <time.h>
#include <map>
#include <mutex>
#include <vector>
#include <algorithm>
static int g_trace = -1;  // -1: read NB_TRACE on first use
struct TraceStat { uint64_t calls, ns; };
static std::map<const char *, TraceStat> g_trace_stats;  // keyed by the address of __func__ (one per entry point)
static std::mutex g_trace_mutex;
static uint64_t now_ns() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (uint64_t)ts.tv_sec * 1000000000ull + ts.tv_nsec;
}
static void trace_print() {
  std::lock_guard<std::mutex> lock(g_trace_mutex);
  if (g_trace_stats.empty()) return;
  std::vector<std::pair<const char *, TraceStat>> rows(g_trace_stats.begin(), g_trace_stats.end());
  std::sort(rows.begin(), rows.end(), [](const std::pair<const char *, TraceStat> &a, const std::pair<const char *, TraceStat> &b) { return a.second.ns > b.second.ns; });
  fprintf(stderr, "[nts_b200 trace, NB_TRACE=%d: %s]\n%-36s %10s %12s %10s\n", g_trace,
          g_trace >= 2 ? "host + GPU time per call (stream synchronised)" : "host time inside each call", "entry point", "calls", "total ms", "mean us");
  for (auto &r : rows)
    fprintf(stderr, "%-36s %10llu %12.3f %10.2f\n", r.first, (unsigned long long)r.second.calls, r.second.ns / 1e6, r.second.ns / 1e3 / r.second.calls);
}
static int trace_level() {
  if (g_trace < 0) {
    const char *e = getenv("NB_TRACE");
    g_trace = e ? atoi(e) : 0;
    if (g_trace > 0) atexit(trace_print);
  }
  return g_trace;
}
bool nb_trace_on() { return trace_level() > 0; }
uint64_t nb_trace_now_ns() { return now_ns(); }
void nb_trace_add(const char *name, uint64_t ns) {
  std::lock_guard<std::mutex> lock(g_trace_mutex);
  TraceStat &s = g_trace_stats[name];
  s.calls++;
  s.ns += ns;
}
void nb_trace_set_level(int level) {
  const bool first = trace_level() <= 0 && level > 0;
  static bool registered = false;
  if (first && !registered && !getenv("NB_TRACE")) { atexit(trace_print); registered = true; }
  g_trace = level;
}
NbTraceScope::NbTraceScope(const char *fn, const nb_ctx *c) : name(nullptr), ctx(c), t0(0) {
  if (trace_level() > 0) { name = fn; t0 = now_ns(); }
}
NbTraceScope::~NbTraceScope() {
  if (!name) return;
  if (g_trace >= 2 && ctx) cudaStreamSynchronize(ctx->stream);
  const uint64_t dt = now_ns() - t0;
  std::lock_guard<std::mutex> lock(g_trace_mutex);
  TraceStat &s = g_trace_stats[name];
  s.calls++;
  s.ns += dt;
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
  if (it != g_mirrors.end() && it->second.size != size) {
    // another allocation now lives at this base (the old one was freed behind our back): the copy is stale
    if (it->second.mirror) { cudaStreamSynchronize(ctx->stream); cudaFree(it->second.mirror); }
    g_mirrors.erase(it);
    it = g_mirrors.end();
  }
  if (it == g_mirrors.end()) {
    void *m = nullptr;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || size > free_b / 2 || cudaMalloc(&m, size) != cudaSuccess) {
      cudaGetLastError();
      return table;   // no room right now: stay zero-copy for this call and try again on the next one
    }
    if (cudaMemcpyAsync(m, (const void *)(uintptr_t)base, size, cudaMemcpyDefault, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) { cudaGetLastError(); cudaFree(m); return table; }
    it = g_mirrors.insert(std::make_pair(key, MirrorEntry{(uintptr_t)base, size, m})).first;
  }
  return (const void *)((const char *)it->second.mirror + ((uintptr_t)attr.devicePointer - it->second.dev_base));
}
// drops (and frees) every device's mirror of the host allocation that contains p; called by nb_free_host and nb_mirror_invalidate
static void mirror_drop(const void *p) {
  if (!p) return;
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return; }
  if (attr.type != cudaMemoryTypeHost || !attr.devicePointer) return;
  const uintptr_t dp = (uintptr_t)attr.devicePointer;
  std::lock_guard<std::mutex> lock(g_mirror_mutex);
  for (auto it = g_mirrors.begin(); it != g_mirrors.end();) {
    if (dp >= it->second.dev_base && dp < it->second.dev_base + it->second.size) {
      if (it->second.mirror) {
        DeviceGuard g(it->first.first);
        cudaDeviceSynchronize();   // kernels reading the mirror have finished
        cudaFree(it->second.mirror);
      }
      it = g_mirrors.erase(it);
    } else ++it;
  }
}
void nb_mirror_host_enable(int on, int adjacency) { if (adjacency) g_mirror_adjacency = on; else g_mirror_tables = on; }


// ---- peer-shareable HBM (CUDA virtual memory management) -------------------------------------------
// Shards of a row-sharded feature table are read by other processes' kernels over NVLink. Mappings made by
// cudaIpcOpenMemHandle translate through small pages: a random-row gather over a multi-GB remote shard ran at ~45 GB/s
// per peer on B200 (TLB-miss bound; tools/shard_bench.py), while the same rows reached through a 2 MB-granular
// cuMemCreate/cuMemMap mapping run at the link rate (~740 GB/s, tools/p2p_probe.cu). So shards are allocated with
// cuMemCreate, exported as POSIX file descriptors and mapped by the peers with cuMemMap.
#include <cuda.h>
#include <unistd.h>
struct VmmEntry { size_t size; CUmemGenericAllocationHandle handle; int fd; };
static std::map<uintptr_t, VmmEntry> g_vmm;
static std::mutex g_vmm_mutex;

template <typename Fn> static bool drv(const char *name, Fn &fn) {
  void *p = nullptr;
  cudaDriverEntryPointQueryResult qr;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &qr) != cudaSuccess || !p) { cudaGetLastError(); return false; }
  fn = (Fn)p;
  return true;
}
struct VmmApi {
  CUresult (*GetGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags);
  CUresult (*Create)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long);
  CUresult (*Release)(CUmemGenericAllocationHandle);
  CUresult (*AddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long);
  CUresult (*AddressFree)(CUdeviceptr, size_t);
  CUresult (*Map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
  CUresult (*Unmap)(CUdeviceptr, size_t);
  CUresult (*SetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t);
  CUresult (*Export)(void *, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
  CUresult (*Import)(CUmemGenericAllocationHandle *, void *, CUmemAllocationHandleType);
  bool ok;
};
static VmmApi *vmm_api() {
  static VmmApi api;
  static bool init = false;
  if (!init) {
    api.ok = drv("cuMemGetAllocationGranularity", api.GetGranularity) && drv("cuMemCreate", api.Create) &&
             drv("cuMemRelease", api.Release) && drv("cuMemAddressReserve", api.AddressReserve) &&
             drv("cuMemAddressFree", api.AddressFree) && drv("cuMemMap", api.Map) && drv("cuMemUnmap", api.Unmap) &&
             drv("cuMemSetAccess", api.SetAccess) && drv("cuMemExportToShareableHandle", api.Export) &&
             drv("cuMemImportFromShareableHandle", api.Import);
    init = true;
  }
  return &api;
}
#define NB_CU(call)                                                                      \
  do {                                                                                   \
    CUresult r_ = (call);                                                                \
    if (r_ != CUDA_SUCCESS) { nb_set_error("%s failed: CUresult %d", #call, (int)r_); return NB_ERR_CUDA; } \
  } while (0)

static CUmemAllocationProp vmm_prop(int device) {
  CUmemAllocationProp prop = {};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return prop;
}
static int vmm_map(VmmApi *a, int device, CUmemGenericAllocationHandle h, size_t size, size_t gran, void **out) {
  CUdeviceptr va = 0;
  NB_CU(a->AddressReserve(&va, size, gran, 0, 0));
  if (a->Map(va, size, 0, h, 0) != CUDA_SUCCESS) { a->AddressFree(va, size); nb_set_error("cuMemMap failed"); return NB_ERR_CUDA; }
  CUmemAccessDesc d = {};
  d.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  d.location.id = device;
  d.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  if (a->SetAccess(va, size, &d, 1) != CUDA_SUCCESS) {
    a->Unmap(va, size); a->AddressFree(va, size);
    nb_set_error("cuMemSetAccess(device %d) failed: no peer path to the owning GPU?", device);
    return NB_ERR_CUDA;
  }
  *out = (void *)va;
  return NB_OK;
}

extern "C" {

int nb_abi_version(void) { return NB_ABI_VERSION; }
const char *nb_last_error(void) { return g_err; }

int nb_trace_dump(void) { trace_print(); return NB_OK; }
int nb_trace_reset(void) { std::lock_guard<std::mutex> lock(g_trace_mutex); g_trace_stats.clear(); return NB_OK; }

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
int nb_free_host(void *p) { if (p) { mirror_drop(p); NB_CUDA(cudaFreeHost(p)); } return NB_OK; }
int nb_mirror_invalidate(const void *host_ptr) { mirror_drop(host_ptr); return NB_OK; }
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

size_t nb_vmm_padded_size(nb_ctx *ctx, size_t bytes) {
  if (!ctx) return 0;
  DeviceGuard g(ctx->device);
  VmmApi *a = vmm_api();
  if (!a->ok) return 0;
  CUmemAllocationProp prop = vmm_prop(ctx->device);
  size_t gran = 0;
  if (a->GetGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || !gran) return 0;
  if (!bytes) bytes = 1;
  return (bytes + gran - 1) / gran * gran;
}

int nb_vmm_alloc(nb_ctx *ctx, size_t bytes, void **dev_ptr_out, int *fd_out) {
  NB_REQUIRE(ctx && dev_ptr_out, NB_ERR_ARG, "nb_vmm_alloc: NULL argument");
  NB_GUARD(ctx);
  NB_CUDA(cudaFree(0));
  VmmApi *a = vmm_api();
  NB_REQUIRE(a->ok, NB_ERR_UNSUPPORTED, "nb_vmm_alloc: the driver lacks the cuMem* virtual memory API");
  CUmemAllocationProp prop = vmm_prop(ctx->device);
  size_t gran = 0;
  NB_CU(a->GetGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  const size_t size = ((bytes ? bytes : 1) + gran - 1) / gran * gran;
  CUmemGenericAllocationHandle h;
  NB_CU(a->Create(&h, size, &prop, 0));
  void *p = nullptr;
  int rc = vmm_map(a, ctx->device, h, size, gran, &p);
  if (rc != NB_OK) { a->Release(h); return rc; }
  int fd = -1;
  if (fd_out) {
    if (a->Export(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) != CUDA_SUCCESS) {
      a->Unmap((CUdeviceptr)p, size); a->AddressFree((CUdeviceptr)p, size); a->Release(h);
      nb_set_error("cuMemExportToShareableHandle failed");
      return NB_ERR_CUDA;
    }
    *fd_out = fd;
  }
  std::lock_guard<std::mutex> lock(g_vmm_mutex);
  g_vmm[(uintptr_t)p] = VmmEntry{size, h, fd};
  *dev_ptr_out = p;
  return NB_OK;
}

int nb_vmm_import(nb_ctx *ctx, int fd, size_t bytes, void **dev_ptr_out) {
  NB_REQUIRE(ctx && dev_ptr_out && fd >= 0, NB_ERR_ARG, "nb_vmm_import: bad argument");
  NB_GUARD(ctx);
  NB_CUDA(cudaFree(0));
  VmmApi *a = vmm_api();
  NB_REQUIRE(a->ok, NB_ERR_UNSUPPORTED, "nb_vmm_import: the driver lacks the cuMem* virtual memory API");
  CUmemAllocationProp prop = vmm_prop(ctx->device);
  size_t gran = 0;
  NB_CU(a->GetGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  const size_t size = ((bytes ? bytes : 1) + gran - 1) / gran * gran;
  CUmemGenericAllocationHandle h;
  NB_CU(a->Import(&h, (void *)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
  void *p = nullptr;
  int rc = vmm_map(a, ctx->device, h, size, gran, &p);
  if (rc != NB_OK) { a->Release(h); return rc; }
  std::lock_guard<std::mutex> lock(g_vmm_mutex);
  g_vmm[(uintptr_t)p] = VmmEntry{size, h, -1};
  *dev_ptr_out = p;
  return NB_OK;
}

int nb_vmm_grant(void *dev_ptr, int device) {
  VmmApi *a = vmm_api();
  NB_REQUIRE(a->ok, NB_ERR_UNSUPPORTED, "nb_vmm_grant: the driver lacks the cuMem* virtual memory API");
  std::lock_guard<std::mutex> lock(g_vmm_mutex);
  auto it = g_vmm.find((uintptr_t)dev_ptr);
  NB_REQUIRE(it != g_vmm.end(), NB_ERR_ARG, "nb_vmm_grant: pointer was not returned by nb_vmm_alloc/import");
  CUmemAccessDesc d = {};
  d.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  d.location.id = device;
  d.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  NB_CU(a->SetAccess((CUdeviceptr)dev_ptr, it->second.size, &d, 1));
  return NB_OK;
}

int nb_vmm_free(void *dev_ptr) {
  if (!dev_ptr) return NB_OK;
  VmmApi *a = vmm_api();
  NB_REQUIRE(a->ok, NB_ERR_UNSUPPORTED, "nb_vmm_free: the driver lacks the cuMem* virtual memory API");
  VmmEntry e;
  {
    std::lock_guard<std::mutex> lock(g_vmm_mutex);
    auto it = g_vmm.find((uintptr_t)dev_ptr);
    NB_REQUIRE(it != g_vmm.end(), NB_ERR_ARG, "nb_vmm_free: pointer was not returned by nb_vmm_alloc/import");
    e = it->second;
    g_vmm.erase(it);
  }
  NB_CU(a->Unmap((CUdeviceptr)dev_ptr, e.size));
  NB_CU(a->AddressFree((CUdeviceptr)dev_ptr, e.size));
  NB_CU(a->Release(e.handle));
  if (e.fd >= 0) close(e.fd);
  return NB_OK;
}

}  // extern "C"
