/* oracle/ref_stubs.cpp -- host-memory stand-ins for the CUDA-side helpers the reference's
 * CPU path touches (TEST INFRASTRUCTURE ONLY). Declarations: cuda/ntsCUDA.hpp:30-71,177-199;
 * the real definitions live in cuda/ntsCUDAGraphOP.cu and need a GPU. The CPU sampler only
 * allocates "pinned" memory (core/FullyRepGraph.hpp:727) and constructs a Cuda_Stream it
 * never launches on (core/ntsFastSampler.hpp:114). */
#include <cstdlib>
#include <cstdint>
#include <cstdio>
#include "cuda/ntsCUDA.hpp"

uint64_t Cuda_Stream::total_sample_num = 0;
uint64_t Cuda_Stream::total_cache_hit = 0;
uint64_t Cuda_Stream::total_transfer_node = 0;
Cuda_Stream::Cuda_Stream() { stream = 0; }
void *cudaMallocPinned(long size_of_bytes) { return calloc(1, size_of_bytes > 0 ? size_of_bytes : 1); }
void *cudaMallocPinnedMulti(long size_of_bytes) { return calloc(1, size_of_bytes > 0 ? size_of_bytes : 1); }
void ntsFreeHost(void *buffer) { free(buffer); }
void *getDevicePointer(void *p) { return p; }

/* Referenced (never called) by the to_gpu=true branches the CPU-only driver does not take:
 * sampCSC::allocate_dev_array_async / copy_data_to_device_async (core/coocsc.hpp:215-300). */
static void nts_stub_unreachable(const char *what) { fprintf(stderr, "oracle stub reached: %s\n", what); abort(); }
void FreeBufferAsync(float *, cudaStream_t) { nts_stub_unreachable("FreeBufferAsync"); }
void FreeEdge(VertexId_CUDA *) { nts_stub_unreachable("FreeEdge"); }
void FreeEdgeAsync(VertexId_CUDA *, cudaStream_t) { nts_stub_unreachable("FreeEdgeAsync"); }
void allocate_gpu_buffer_async(float **, int, cudaStream_t) { nts_stub_unreachable("allocate_gpu_buffer_async"); }
void allocate_gpu_edge(VertexId_CUDA **, int) { nts_stub_unreachable("allocate_gpu_edge"); }
void allocate_gpu_edge_async(VertexId_CUDA **, int, cudaStream_t) { nts_stub_unreachable("allocate_gpu_edge_async"); }
void *cudaMallocGPU(long) { nts_stub_unreachable("cudaMallocGPU"); return 0; }
void move_bytes_in(void *, void *, long, bool) { nts_stub_unreachable("move_bytes_in"); }
void move_bytes_in_async(void *, void *, long, cudaStream_t) { nts_stub_unreachable("move_bytes_in_async"); }
/* comm/network.cpp frees its (never allocated here) pinned message buffers */
extern "C" cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
