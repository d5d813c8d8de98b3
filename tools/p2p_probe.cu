// p2p_probe.cu -- why is a random-row peer gather slow over a large remote footprint?
// Single process, GPUs 0 (reader) and 1 (owner). Random 512-byte rows are read from GPU 1's memory by a kernel on GPU 0.
// Allocation kinds: (A) cudaMalloc + cudaDeviceEnablePeerAccess, (B) cuMemCreate (granularity = recommended) + cuMemMap +
// cuMemSetAccess, VA aligned to 2 MB, (C) the same with size and VA aligned to 512 MB. Footprint sweep per kind.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/p2p_probe tools/p2p_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
#define CU(x) do { CUresult e = (x); if (e != CUDA_SUCCESS) { const char *s; cuGetErrorString(e, &s); printf("CU %s at %d\n", s, __LINE__); exit(1); } } while (0)

template <int RPW>
__global__ void __launch_bounds__(256) k_gather(float4 *__restrict__ out, const float4 *__restrict__ table, const uint32_t *__restrict__ ids, uint32_t n) {
  const unsigned lane = threadIdx.x & 31;
  const unsigned warp = (blockIdx.x * 256 + threadIdx.x) >> 5, warps = (gridDim.x * 256) >> 5;
  for (unsigned i = warp * RPW; i < n; i += warps * RPW) {
    float4 x[RPW];
#pragma unroll
    for (int r = 0; r < RPW; r++)
      if (i + r < n) x[r] = __ldg(table + (uint64_t)ids[i + r] * 32 + lane);
#pragma unroll
    for (int r = 0; r < RPW; r++)
      if (i + r < n) out[(uint64_t)(i + r) * 32 + lane] = x[r];
  }
}

__global__ void k_ids(uint32_t *ids, uint32_t n, uint32_t rows, uint64_t seed) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint64_t z = seed + (uint64_t)i * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    ids[i] = (uint32_t)(z % rows);
  }
}

static float time_gather(int rpw, float4 *out, const float4 *table, const uint32_t *ids, uint32_t n) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  const int grid = 148 * 8;
  for (int it = 0; it < 2; it++) {
    if (rpw == 1) k_gather<1><<<grid, 256>>>(out, table, ids, n); else k_gather<4><<<grid, 256>>>(out, table, ids, n);
  }
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int it = 0; it < 5; it++) {
    if (rpw == 1) k_gather<1><<<grid, 256>>>(out, table, ids, n); else k_gather<4><<<grid, 256>>>(out, table, ids, n);
  }
  CK(cudaEventRecord(b)); CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  return ms / 5;
}

struct VmmAlloc { CUdeviceptr va; size_t size; CUmemGenericAllocationHandle h; };

static VmmAlloc vmm_alloc(int owner, size_t bytes, size_t align, int n_access, const int *access_devs) {
  CUmemAllocationProp prop = {};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = owner;
  size_t gran; CU(cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  if (align < gran) align = gran;
  VmmAlloc a; a.size = (bytes + align - 1) / align * align;
  CU(cuMemCreate(&a.h, a.size, &prop, 0));
  CU(cuMemAddressReserve(&a.va, a.size, align, 0, 0));
  CU(cuMemMap(a.va, a.size, 0, a.h, 0));
  CUmemAccessDesc d[8];
  for (int i = 0; i < n_access; i++) { d[i].location.type = CU_MEM_LOCATION_TYPE_DEVICE; d[i].location.id = access_devs[i]; d[i].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE; }
  CU(cuMemSetAccess(a.va, a.size, d, n_access));
  return a;
}
static void vmm_free(VmmAlloc &a) { CU(cuMemUnmap(a.va, a.size)); CU(cuMemAddressFree(a.va, a.size)); CU(cuMemRelease(a.h)); }

int main(int argc, char **argv) {
  int ndev; CK(cudaGetDeviceCount(&ndev));
  if (ndev < 2) { printf("need 2 GPUs\n"); return 0; }
  CK(cudaSetDevice(1)); CK(cudaFree(0));
  CK(cudaSetDevice(0)); CK(cudaFree(0));
  CK(cudaDeviceEnablePeerAccess(1, 0));
  const uint32_t N = 400000;
  float4 *out; uint32_t *ids;
  CK(cudaMalloc(&out, (size_t)N * 512)); CK(cudaMalloc(&ids, N * 4));
  size_t gran_min, gran_rec; CUmemAllocationProp prop = {};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = 1;
  CU(cuMemGetAllocationGranularity(&gran_min, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM));
  CU(cuMemGetAllocationGranularity(&gran_rec, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  printf("VMM granularity min %zu recommended %zu\n", gran_min, gran_rec);
  const double gbs[] = {0.25, 1, 2, 4, 8, 14};
  const int devs[2] = {0, 1};
  for (double gb : gbs) {
    const size_t bytes = (size_t)(gb * (1ull << 30));
    const uint32_t rows = (uint32_t)(bytes / 512);
    k_ids<<<256, 256>>>(ids, N, rows, 12345 + (uint64_t)bytes);
    CK(cudaDeviceSynchronize());
    // local baseline
    { float4 *t; CK(cudaMalloc(&t, bytes)); CK(cudaMemset(t, 1, bytes));
      float m1 = time_gather(1, out, t, ids, N), m4 = time_gather(4, out, t, ids, N);
      printf("foot %5.2f GB  local cudaMalloc      : rpw1 %7.1f GB/s  rpw4 %7.1f GB/s\n", gb, N * 512.0 / m1 / 1e6, N * 512.0 / m4 / 1e6);
      CK(cudaFree(t)); }
    // A: cudaMalloc on 1
    { float4 *t; CK(cudaSetDevice(1)); CK(cudaMalloc(&t, bytes)); CK(cudaMemset(t, 1, bytes)); CK(cudaDeviceSynchronize()); CK(cudaSetDevice(0));
      float m1 = time_gather(1, out, t, ids, N), m4 = time_gather(4, out, t, ids, N);
      printf("foot %5.2f GB  peer  cudaMalloc      : rpw1 %7.1f GB/s  rpw4 %7.1f GB/s\n", gb, N * 512.0 / m1 / 1e6, N * 512.0 / m4 / 1e6);
      CK(cudaSetDevice(1)); CK(cudaFree(t)); CK(cudaSetDevice(0)); }
    // B / C: VMM
    for (size_t align : {(size_t)2 << 20, (size_t)512 << 20}) {
      VmmAlloc a = vmm_alloc(1, bytes, align, 2, devs);
      CK(cudaMemset((void *)a.va, 1, bytes)); CK(cudaDeviceSynchronize());
      float m1 = time_gather(1, out, (const float4 *)a.va, ids, N), m4 = time_gather(4, out, (const float4 *)a.va, ids, N);
      printf("foot %5.2f GB  peer  VMM align %4zuMB: rpw1 %7.1f GB/s  rpw4 %7.1f GB/s\n", gb, align >> 20, N * 512.0 / m1 / 1e6, N * 512.0 / m4 / 1e6);
      vmm_free(a);
    }
    fflush(stdout);
  }
  return 0;
}
