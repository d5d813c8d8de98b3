/* ntsCUDA.hpp -- header-only adaptor with the reference's host<->CUDA boundary (cuda/ntsCUDA.hpp of
 * AiX-im/Sample-based-GNN) implemented on top of the C ABI of libnts_b200.so (include/nts_b200.h).
 *
 * Put this directory before the reference's on the include path (-I sample-based-gnn_b200/host
 * -I sample-based-gnn_b200/host/cuda) and link -lnts_b200 instead of the reference's cuda_propagate
 * static library: core/*.hpp and toolkits/*.hpp compile unchanged (tests/test_adaptor_compile.py
 * does exactly that with the reference's own headers). Same free functions (:30-71), deviceCSC
 * (:73-121), NCCL helpers (:123-175) and class Cuda_Stream (:177-595) with identical member names,
 * signatures and public data members.
 *
 * Error convention: like the reference (cuda/ntsCUDAGraphOP.cu:21-60) a failing call prints and
 * exit(1)s -- on top of the status codes of the C ABI. Methods of the reference that belong to its
 * full-graph (non-sampled) engine are declared for source compatibility and fail loudly if called:
 * they are outside the sampled hot path this library replaces.
 */
/* VertexId_CUDA and the launch constants core/ reads: taken from the reference tree's own cuda/cuda_type.h when it is on the
 * include path (it is, when the reference is being built against this adaptor); otherwise the same names are defined here */
#if __has_include("cuda/cuda_type.h")
#include "cuda/cuda_type.h"
#else
#ifndef CUDA_TYPE_H
#define CUDA_TYPE_H
#include <stdint.h>
typedef uint32_t VertexId_CUDA;
static const int WARP_SIZE = 32, CUDA_NUM_THREADS = 512, CUDA_NUM_BLOCKS = 128, CUDA_NUM_THREADS_SOFTMAX = 32, CUDA_NUM_BLOCKS_SOFTMAX = 512;
#endif
#endif
#define CUDA_ENABLE 1
#include <cuda_runtime.h>
#if __has_include(<nccl.h>)
#include <nccl.h>
#define NTS_B200_HAVE_NCCL 1
#else
typedef struct ncclComm *ncclComm_t;
#endif

#ifndef TEST_HPP
#define TEST_HPP
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <fstream>
#include <iostream>
#include <vector>

#include "nts_b200.h"

/* the reference's handle types appear as public members of Cuda_Stream; kept as opaque pointers */
#ifndef CUSPARSE_H_
typedef struct cusparseContext *cusparseHandle_t;
#endif
#ifndef CUBLAS_API_H_
typedef struct cublasContext *cublasHandle_t;
#endif

enum graph_type { CSR, CSC, PAIR };
enum weight_type { NULL_TYPE, SCALA_TYPE, TENSOR_TYPE };

#define NTS_B200_CHECK(call)                                                                         \
  do {                                                                                               \
    int rc__ = (call);                                                                               \
    if (rc__ != 0) {                                                                                 \
      fprintf(stderr, "libnts_b200 error %d at %s:%d: %s\n", rc__, __FILE__, __LINE__, nb_last_error()); \
      exit(1);                                                                                       \
    }                                                                                                \
  } while (0)
#define NTS_B200_UNSUPPORTED(name)                                                                   \
  do {                                                                                               \
    fprintf(stderr, "Cuda_Stream::%s belongs to the full-graph engine and is not provided by libnts_b200\n", name); \
    exit(1);                                                                                         \
  } while (0)

/* ---- free functions, cuda/ntsCUDA.hpp:30-71 --------------------------------------------------- */
inline void ntsFreeHost(void *buffer) { NTS_B200_CHECK(nb_free_host(buffer)); }
inline void *cudaMallocPinned(long size_of_bytes) { void *p = 0; NTS_B200_CHECK(nb_malloc_pinned((size_t)size_of_bytes, &p)); return p; }
inline void *cudaMallocPinnedMulti(long size_of_bytes) { return cudaMallocPinned(size_of_bytes); }
inline void *cudaMallocGPU(long size_of_bytes) { void *p = 0; NTS_B200_CHECK(nb_malloc_device((size_t)size_of_bytes, &p)); return p; }
inline void *cudaMallocZero(long size_of_bytes) { void *p = cudaMallocGPU(size_of_bytes); cudaMemset(p, 0, (size_t)size_of_bytes); return p; }
inline void *getDevicePointer(void *host_data_to_device) { void *p = 0; NTS_B200_CHECK(nb_device_pointer(host_data_to_device, &p)); return p; }
inline void cudaSetMemAsync(void *mem, int value, size_t size, cudaStream_t stream) { cudaMemsetAsync(mem, value, size, stream); }
inline void cudaSetUsingDevice(int device_id) { if (cudaSetDevice(device_id) != cudaSuccess) { fprintf(stderr, "cudaSetDevice(%d) failed\n", device_id); exit(1); } }
inline void move_bytes_in(void *d_pointer, void *h_pointer, long bytes, bool sync = true) {
  cudaMemcpy(d_pointer, h_pointer, (size_t)bytes, cudaMemcpyHostToDevice); if (sync) cudaDeviceSynchronize(); }
inline void move_bytes_in_async(void *d_pointer, void *h_pointer, long bytes, cudaStream_t cs) { cudaMemcpyAsync(d_pointer, h_pointer, (size_t)bytes, cudaMemcpyHostToDevice, cs); }
inline void move_bytes_in_async_check(void *d_pointer, void *h_pointer, long bytes, cudaStream_t cs) { move_bytes_in_async(d_pointer, h_pointer, bytes, cs); }
inline void move_bytes_out(void *h_pointer, void *d_pointer, long bytes, bool sync = true) {
  cudaMemcpy(h_pointer, d_pointer, (size_t)bytes, cudaMemcpyDeviceToHost); if (sync) cudaDeviceSynchronize(); }
inline void move_bytes_out_async(void *h_pointer, void *d_pointer, long bytes, cudaStream_t cs) { cudaMemcpyAsync(h_pointer, d_pointer, (size_t)bytes, cudaMemcpyDeviceToHost, cs); }
inline void move_result_out(float *output, float *input, int src, int dst, int feature_size, bool sync = true) {
  move_bytes_out(output, input, (long)(dst - src) * feature_size * (long)sizeof(float), sync); }
inline void move_data_in(float *d_pointer, float *h_pointer, int start, int end, int feature_size, bool sync = true) {
  move_bytes_in(d_pointer, h_pointer, (long)(end - start) * feature_size * (long)sizeof(float), sync); }
inline void move_edge_in(VertexId_CUDA *d_pointer, VertexId_CUDA *h_pointer, VertexId_CUDA start, VertexId_CUDA end, int feature_size, bool sync = true) {
  move_bytes_in(d_pointer, h_pointer, (long)(end - start) * feature_size * (long)sizeof(VertexId_CUDA), sync); }
/* sizes are element counts; computed in size_t (the reference's `int size * sizeof` overflows past 2 GiB) */
inline void allocate_gpu_buffer(float **input, int size) { *input = (float *)cudaMallocGPU((long)size * (long)sizeof(float)); }
inline void allocate_gpu_edge(VertexId_CUDA **input, int size) { *input = (VertexId_CUDA *)cudaMallocGPU((long)size * (long)sizeof(VertexId_CUDA)); }
inline void free_gpu_mem_async(void *mem, cudaStream_t cs) { cudaFreeAsync(mem, cs); }
inline void allocate_gpu_buffer_async(float **input, int size, cudaStream_t cs) { cudaMallocAsync((void **)input, (size_t)size * sizeof(float), cs); }
inline void allocate_gpu_edge_async(VertexId_CUDA **input, int size, cudaStream_t cs) { cudaMallocAsync((void **)input, (size_t)size * sizeof(VertexId_CUDA), cs); }
inline void FreeBuffer(float *buffer) { NTS_B200_CHECK(nb_free_device(buffer)); }
inline void FreeEdge(VertexId_CUDA *buffer) { NTS_B200_CHECK(nb_free_device(buffer)); }
inline void FreeBufferAsync(float *buffer, cudaStream_t cs) { cudaFreeAsync(buffer, cs); }
inline void FreeEdgeAsync(VertexId_CUDA *buffer, cudaStream_t cs) { cudaFreeAsync(buffer, cs); }
inline void zero_buffer(float *buffer, int size) { cudaMemset(buffer, 0, (size_t)size * sizeof(float)); }
inline void CUDA_DEVICE_SYNCHRONIZE() { cudaDeviceSynchronize(); }
inline void ResetDevice() { cudaDeviceReset(); }
inline void aggregate_comm_result(float *, float *, int, int, int, bool = true) { NTS_B200_UNSUPPORTED("aggregate_comm_result"); }

/* ---- deviceCSC, cuda/ntsCUDA.hpp:73-121 -------------------------------------------------------- */
class deviceCSC {
public:
  VertexId_CUDA *column_offset;
  VertexId_CUDA *row_indices;
  VertexId_CUDA *mirror_index;
  VertexId_CUDA v_size;
  VertexId_CUDA e_size;
  VertexId_CUDA mirror_size;
  bool require_mirror = false;
  deviceCSC() { column_offset = NULL; row_indices = NULL; }
  void init(VertexId_CUDA v_size_, VertexId_CUDA e_size_, bool require_mirror_ = false, VertexId_CUDA mirror_size_ = 0) {
    v_size = v_size_; e_size = e_size_; require_mirror = false;
    column_offset = (VertexId_CUDA *)cudaMallocGPU(((long)v_size_ + 1) * (long)sizeof(VertexId_CUDA));
    row_indices = (VertexId_CUDA *)cudaMallocGPU((long)e_size_ * (long)sizeof(VertexId_CUDA));
    if (require_mirror_) {
      require_mirror = require_mirror_; mirror_size = mirror_size_;
      mirror_index = (VertexId_CUDA *)cudaMallocGPU((long)mirror_size_ * (long)sizeof(VertexId_CUDA));
    }
  }
  void load_from_host(VertexId_CUDA *h_column_offset, VertexId_CUDA *h_row_indices, VertexId_CUDA *h_mirror_index) {
    move_bytes_in(column_offset, h_column_offset, ((long)v_size + 1) * (long)sizeof(VertexId_CUDA));
    move_bytes_in(row_indices, h_row_indices, (long)e_size * (long)sizeof(VertexId_CUDA));
    move_bytes_in(mirror_index, h_mirror_index, (long)mirror_size * (long)sizeof(VertexId_CUDA));
  }
  void load_from_host(VertexId_CUDA *h_column_offset, VertexId_CUDA *h_row_indices) {
    move_bytes_in(column_offset, h_column_offset, ((long)v_size + 1) * (long)sizeof(VertexId_CUDA));
    move_bytes_in(row_indices, h_row_indices, (long)e_size * (long)sizeof(VertexId_CUDA));
  }
  void release() { FreeEdge(column_offset); FreeEdge(row_indices); if (require_mirror) FreeEdge(mirror_index); }
  ~deviceCSC() {}
};

/* ---- NCCL helpers + NCCL_Communicator, cuda/ntsCUDA.hpp:123-175 (cuda/ntsCUDAGraphOP.cu:173-200) ---- */
#ifdef NTS_B200_HAVE_NCCL
#define NTS_B200_NCCL(call) do { ncclResult_t r__ = (call); if (r__ != ncclSuccess) { fprintf(stderr, "NCCL error %s at %s:%d\n", ncclGetErrorString(r__), __FILE__, __LINE__); exit(1); } } while (0)
inline void destroyNCCLComm(ncclComm_t comm) { ncclCommDestroy(comm); }
inline void initNCCLComm(ncclComm_t *comms, int nDev, int *devs) { NTS_B200_NCCL(ncclCommInitAll(comms, nDev, devs)); }
inline void allReduceNCCL(void *send_buffer, void *recv_buffer, size_t element_num, ncclComm_t comm, cudaStream_t cudaStream, int) {
  NTS_B200_NCCL(ncclAllReduce(send_buffer, recv_buffer, element_num, ncclFloat, ncclSum, comm, cudaStream)); }
inline void broadcastNCCL(void *send_buffer, size_t element_num, ncclComm_t comm, cudaStream_t cudaStream, int root) {
  NTS_B200_NCCL(ncclBcast(send_buffer, element_num, ncclFloat, root, comm, cudaStream)); }
/* the reference passes the same base pointer as send and receive buffer on every rank (ntsDataloador.hpp:762); the intent --
 * all-gather of equal slices of one buffer -- is what is implemented: rank r contributes slice r of recv_buffer */
inline void allGatherNCCL(void *send_buffer, void *recv_buffer, size_t element_num, ncclComm_t comm, cudaStream_t cudaStream, int) {
  int rank = 0; NTS_B200_NCCL(ncclCommUserRank(comm, &rank));
  const void *mine = send_buffer == recv_buffer ? (const void *)((const float *)recv_buffer + (size_t)rank * element_num) : send_buffer;
  NTS_B200_NCCL(ncclAllGather(mine, recv_buffer, element_num, ncclFloat, comm, cudaStream)); }
#else
inline void destroyNCCLComm(ncclComm_t) {}
inline void initNCCLComm(ncclComm_t *, int, int *) { fprintf(stderr, "built without <nccl.h>\n"); exit(1); }
inline void allReduceNCCL(void *, void *, size_t, ncclComm_t, cudaStream_t, int) { fprintf(stderr, "built without <nccl.h>\n"); exit(1); }
inline void broadcastNCCL(void *, size_t, ncclComm_t, cudaStream_t, int) { fprintf(stderr, "built without <nccl.h>\n"); exit(1); }
inline void allGatherNCCL(void *, void *, size_t, ncclComm_t, cudaStream_t, int) { fprintf(stderr, "built without <nccl.h>\n"); exit(1); }
#endif

class NCCL_Communicator {
private:
  int device_num;
  int root;
  ncclComm_t *ncclComms;
  void initAllNCCLComm(int nDev, int *devs) { ncclComms = new ncclComm_t[nDev]; initNCCLComm(ncclComms, nDev, devs); }
public:
  NCCL_Communicator(int device_num_, int *devs, int root_ = 0) { this->device_num = device_num_; this->root = root_; initAllNCCLComm(device_num_, devs); }
  ~NCCL_Communicator() { for (int i = 0; i < device_num; i++) destroyNCCLComm(ncclComms[i]); delete[] ncclComms; }
  void AllReduce(int device_id, void *send_buffer, void *recv_buffer, size_t element_num, cudaStream_t cudaStream) {
    allReduceNCCL(send_buffer, recv_buffer, element_num, ncclComms[device_id], cudaStream, device_id); }
  void Bcast(int device_id, void *send_buffer, size_t element_num, cudaStream_t cudaStream) {
    broadcastNCCL(send_buffer, element_num, ncclComms[device_id], cudaStream, root); }
  void AllGather(int device_id, void *send_buffer, size_t element_num, cudaStream_t cudaStream) {
    allGatherNCCL(send_buffer, send_buffer, element_num, ncclComms[device_id], cudaStream, device_id); }
};

/* ---- class Cuda_Stream, cuda/ntsCUDA.hpp:177-595 ------------------------------------------------ */
class Cuda_Stream {
public:
  double cpu_inclusiveTime = 0.0;
  double inclusiveTime = 0;
  static inline uint64_t total_sample_num = 0;
  static inline uint64_t total_cache_hit = 0;
  static inline uint64_t total_transfer_node = 0;

  cudaStream_t stream;
  cusparseHandle_t sparse_handle = NULL; /* unused: aggregation is our own segment reduce, not cuSPARSE */
  cublasHandle_t blas_handle = NULL;
  void *cuda_buffer = NULL;
  size_t cuda_buffer_size = 0;
  unsigned char *cpu_buffer = NULL;
  size_t cpu_buffer_size = 0;
  /* additions */
  nb_ctx *ctx = NULL;
  uint64_t rng_seed = 0x5EED0004ull; /* Philox key; the reference reseeds from std::random_device per launch */
  uint64_t rng_counter = 0;

  Cuda_Stream() {
    int dev = 0;
    cudaGetDevice(&dev);
    NTS_B200_CHECK(nb_ctx_create(dev, NULL, 0, &ctx));
    stream = (cudaStream_t)nb_ctx_stream(ctx);
  }
  void destory_Stream() { if (ctx) { NTS_B200_CHECK(nb_ctx_destroy(ctx)); ctx = NULL; } }
  cudaStream_t getStream() { return stream; }
  void setNewStream(cudaStream_t cudaStream) { NTS_B200_CHECK(nb_ctx_set_stream(ctx, cudaStream)); stream = cudaStream; }
  void CUDA_DEVICE_SYNCHRONIZE() { NTS_B200_CHECK(nb_ctx_synchronize(ctx)); }
  void CUDA_SYNCHRONIZE_ALL() { cudaDeviceSynchronize(); }

  void move_bytes_out(VertexId_CUDA *h_pointer, VertexId_CUDA *d_pointer, int size) {
    NTS_B200_CHECK(nb_memcpy_d2h(ctx, h_pointer, d_pointer, (size_t)size * sizeof(VertexId_CUDA), 1)); }
  void move_result_out(float *output, float *input, VertexId_CUDA src, VertexId_CUDA dst, int feature_size, bool sync = true) {
    NTS_B200_CHECK(nb_memcpy_d2h(ctx, output, input, (size_t)(dst - src) * feature_size * sizeof(float), sync)); }
  void move_data_in(float *d_pointer, float *h_pointer, VertexId_CUDA start, VertexId_CUDA end, int feature_size, bool sync = true) {
    NTS_B200_CHECK(nb_memcpy_h2d(ctx, d_pointer, h_pointer, (size_t)(end - start) * feature_size * sizeof(float), sync)); }
  void move_edge_in(VertexId_CUDA *d_pointer, VertexId_CUDA *h_pointer, VertexId_CUDA start, VertexId_CUDA end, int feature_size, bool sync = true) {
    NTS_B200_CHECK(nb_memcpy_h2d(ctx, d_pointer, h_pointer, (size_t)(end - start) * feature_size * sizeof(VertexId_CUDA), sync)); }
  void aggregate_comm_result(float *, float *, VertexId_CUDA, int, int, bool = true) { NTS_B200_UNSUPPORTED("aggregate_comm_result"); }
  void deSerializeToGPU(float *, float *, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, bool) { NTS_B200_UNSUPPORTED("deSerializeToGPU"); }
  void aggregate_comm_result_debug(float *, float *, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, bool) { NTS_B200_UNSUPPORTED("aggregate_comm_result_debug"); }

  /* -- aggregation: every variant of the reference maps onto the two segment reductions (+ the CSC push) */
  void Gather_By_Dst_From_Src(float *input, float *output, float *weight_forward, VertexId_CUDA *row_indices, VertexId_CUDA *column_offset,
                              VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA batch_size,
                              VertexId_CUDA feature_size, bool with_weight = false, bool = false) {
    NTS_B200_CHECK(nb_aggregate_csc_fwd(ctx, input, output, with_weight ? weight_forward : NULL, row_indices, column_offset, batch_size, 0, feature_size)); }
  void Gather_By_Dst_From_Src_Spmm(float *input, float *output, float *weight_forward, VertexId_CUDA *row_indices, VertexId_CUDA *column_offset,
                                   VertexId_CUDA column_num, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA,
                                   VertexId_CUDA batch_size, VertexId_CUDA feature_size, bool with_weight = false, bool = false) {
    NTS_B200_CHECK(nb_aggregate_csc_fwd(ctx, input, output, with_weight ? weight_forward : NULL, row_indices, column_offset, batch_size, column_num, feature_size)); }
  void Gather_By_Dst_From_Src_Optim(float *input, float *output, float *weight_forward, VertexId_CUDA *row_indices, VertexId_CUDA *column_offset,
                                    VertexId_CUDA a, VertexId_CUDA b, VertexId_CUDA c, VertexId_CUDA d, VertexId_CUDA edges, VertexId_CUDA batch_size,
                                    VertexId_CUDA feature_size, bool with_weight = false, bool tensor_weight = false) {
    Gather_By_Dst_From_Src(input, output, weight_forward, row_indices, column_offset, a, b, c, d, edges, batch_size, feature_size, with_weight, tensor_weight); }
  void Gather_By_Dst_From_Src_with_cache(float *, float *, float *, VertexId_CUDA *, VertexId_CUDA *, VertexId_CUDA *, VertexId_CUDA *, VertexId_CUDA,
                                         VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, bool = false, bool = false) {
    NTS_B200_UNSUPPORTED("Gather_By_Dst_From_Src_with_cache"); }
  void Push_From_Dst_To_Src(float *input, float *output, float *weight_forward, VertexId_CUDA *row_indices, VertexId_CUDA *column_offset,
                            VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, bool = false, bool = false) {
    (void)input; (void)output; (void)weight_forward; (void)row_indices; (void)column_offset;
    /* needs the source count, which only the _Spmm overload carries */
    NTS_B200_UNSUPPORTED("Push_From_Dst_To_Src (use Push_From_Dst_To_Src_Spmm, which carries column_num)"); }
  void Push_From_Dst_To_Src_Spmm(float *input, float *output, float *weight_forward, VertexId_CUDA *row_indices, VertexId_CUDA *column_offset,
                                 int column_num, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA batch_size,
                                 VertexId_CUDA feature_size, bool with_weight = false, bool = false) {
    NTS_B200_CHECK(nb_aggregate_push_bwd(ctx, input, output, with_weight ? weight_forward : NULL, row_indices, column_offset, batch_size, (uint32_t)column_num, feature_size)); }
  void Gather_By_Src_From_Dst(float *input, float *output, float *weight_backward, VertexId_CUDA *row_offset, VertexId_CUDA *column_indices,
                              VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA batch_size,
                              VertexId_CUDA feature_size, bool with_weight = false, bool = false) {
    NTS_B200_CHECK(nb_aggregate_csr_bwd(ctx, input, output, with_weight ? weight_backward : NULL, row_offset, column_indices, batch_size, 0, feature_size)); }
  void Gather_By_Src_From_Dst_Optim(float *input, float *output, float *weight_backward, VertexId_CUDA *row_offset, VertexId_CUDA *column_indices,
                                    VertexId_CUDA a, VertexId_CUDA b, VertexId_CUDA c, VertexId_CUDA d, VertexId_CUDA edges, VertexId_CUDA batch_size,
                                    VertexId_CUDA feature_size, bool with_weight = false, bool tensor_weight = false) {
    Gather_By_Src_From_Dst(input, output, weight_backward, row_offset, column_indices, a, b, c, d, edges, batch_size, feature_size, with_weight, tensor_weight); }
  void Gather_By_Src_From_Dst_Spmm(float *input, float *output, float *weight_backward, VertexId_CUDA *row_offset, VertexId_CUDA *column_indices,
                                   VertexId_CUDA column_num, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA,
                                   VertexId_CUDA batch_size, VertexId_CUDA feature_size, bool with_weight = false, bool = false) {
    NTS_B200_CHECK(nb_aggregate_csr_bwd(ctx, input, output, with_weight ? weight_backward : NULL, row_offset, column_indices, batch_size, column_num, feature_size)); }

  /* -- full-graph edge ops (mirror / un-mapped variants): not on the sampled path */
  void Scatter_Src_Mirror_to_Msg(float *, float *, VertexId_CUDA *, VertexId_CUDA *, VertexId_CUDA *, VertexId_CUDA, VertexId_CUDA) { NTS_B200_UNSUPPORTED("Scatter_Src_Mirror_to_Msg"); }
  void Scatter_Src_to_Msg(float *, float *, VertexId_CUDA *, VertexId_CUDA *, VertexId_CUDA, VertexId_CUDA) { NTS_B200_UNSUPPORTED("Scatter_Src_to_Msg"); }
  void Gather_Msg_To_Src_Mirror(float *, float *, VertexId_CUDA *, VertexId_CUDA *, VertexId_CUDA *, VertexId_CUDA, VertexId_CUDA) { NTS_B200_UNSUPPORTED("Gather_Msg_To_Src_Mirror"); }
  void Gather_Msg_To_Src(float *, float *, VertexId_CUDA *, VertexId_CUDA *, VertexId_CUDA, VertexId_CUDA) { NTS_B200_UNSUPPORTED("Gather_Msg_To_Src"); }
  void Edge_Softmax_Forward_Block(float *, float *, float *, VertexId_CUDA *, VertexId_CUDA *, VertexId_CUDA, VertexId_CUDA) { NTS_B200_UNSUPPORTED("Edge_Softmax_Forward_Block (replaced by Edge_Softmax_Forward_Norm_Block)"); }
  void Gather_By_Dst_From_Message(float *, float *, VertexId_CUDA *, VertexId_CUDA *, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, bool = false, bool = false) { NTS_B200_UNSUPPORTED("Gather_By_Dst_From_Message"); }
  void Scatter_Grad_Back_To_Message(float *, float *, VertexId_CUDA *, VertexId_CUDA *, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA, bool = true) { NTS_B200_UNSUPPORTED("Scatter_Grad_Back_To_Message"); }
  void Scatter_Src_to_Msg_Map(float *, float *, VertexId_CUDA *, VertexId_CUDA *, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA *) { NTS_B200_UNSUPPORTED("Scatter_Src_to_Msg_Map"); }
  void Gather_Msg_To_Src_Map(float *, float *, VertexId_CUDA *, VertexId_CUDA *, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA *) { NTS_B200_UNSUPPORTED("Gather_Msg_To_Src_Map"); }
  void Scatter_Dst_to_Msg_Map(float *, float *, VertexId_CUDA *, VertexId_CUDA *, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA *) { NTS_B200_UNSUPPORTED("Scatter_Dst_to_Msg_Map"); }
  void Gather_Msg_to_Dst_Map(float *, float *, VertexId_CUDA *, VertexId_CUDA *, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA *) { NTS_B200_UNSUPPORTED("Gather_Msg_to_Dst_Map"); }

  /* -- GAT edge ops on the sampled path (BatchGPU* ops, core/ntsPushdownGraphOp.hpp:490-747) */
  void Scatter_Dst_to_Msg(float *message, float *dst_feature, VertexId_CUDA *, VertexId_CUDA *column_offset, VertexId_CUDA batch_size, VertexId_CUDA feature_size) {
    NTS_B200_CHECK(nb_scatter_dst_to_msg(ctx, message, dst_feature, column_offset, batch_size, feature_size)); }
  void Gather_Msg_to_Dst(float *dst_feature, float *message, VertexId_CUDA *, VertexId_CUDA *column_offset, VertexId_CUDA batch_size, VertexId_CUDA feature_size) {
    NTS_B200_CHECK(nb_gather_msg_to_dst(ctx, dst_feature, message, column_offset, batch_size, feature_size)); }
  void Edge_Softmax_Forward_Norm_Block(float *msg_output, float *msg_input, float *msg_cached, VertexId_CUDA *, VertexId_CUDA *column_offset, VertexId_CUDA batch_size, VertexId_CUDA feature_size) {
    if (feature_size != 1) NTS_B200_UNSUPPORTED("Edge_Softmax_Forward_Norm_Block with feature_size != 1");
    NTS_B200_CHECK(nb_edge_softmax_fwd(ctx, msg_output, msg_input, msg_cached, column_offset, batch_size)); }
  void Edge_Softmax_Backward_Block(float *msg_input_grad, float *msg_output_grad, float *msg_cached, VertexId_CUDA *, VertexId_CUDA *column_offset, VertexId_CUDA batch_size, VertexId_CUDA feature_size) {
    if (feature_size != 1) NTS_B200_UNSUPPORTED("Edge_Softmax_Backward_Block with feature_size != 1");
    NTS_B200_CHECK(nb_edge_softmax_bwd(ctx, msg_input_grad, msg_output_grad, msg_cached, column_offset, batch_size)); }
  void Scatter_Src_Dst_to_Msg(float *message, float *src_mirror_feature, VertexId_CUDA *row_indices, VertexId_CUDA *column_offset, VertexId_CUDA batch_size,
                              VertexId_CUDA feature_size, VertexId_CUDA *dst_to_local) {
    NTS_B200_CHECK(nb_scatter_src_dst_to_msg(ctx, message, src_mirror_feature, row_indices, column_offset, batch_size, feature_size, dst_to_local)); }
  /* the caller's output tensor is pre-zeroed (NewKeyTensor -> torch::zeros, core/NtsScheduler.hpp:398-414), as the reference's atomics
   * require; n_src is not part of this signature, so the accumulation relies on that contract (n_src == 0 skips our own memset) */
  void Gather_Msg_To_Src_Dst(float *src_mirror_feature, float *message, VertexId_CUDA *row_indices, VertexId_CUDA *column_offset, VertexId_CUDA batch_size,
                             VertexId_CUDA feature_size, VertexId_CUDA *dst_to_local) {
    NTS_B200_CHECK(nb_gather_msg_to_src_dst(ctx, src_mirror_feature, message, row_indices, column_offset, batch_size, 0, feature_size, dst_to_local)); }

  /* -- sampling stages (core/FullyRepGraph.hpp:326-524) */
  void sample_processing_get_co_gpu(VertexId_CUDA *dst, VertexId_CUDA *local_column_offset, VertexId_CUDA *global_column_offset, VertexId_CUDA dst_size,
                                    VertexId_CUDA *, VertexId_CUDA, VertexId_CUDA *, VertexId_CUDA *, VertexId_CUDA fanout, VertexId_CUDA &edge_size) {
    NTS_B200_CHECK(nb_sample_count(ctx, dst, local_column_offset, global_column_offset, dst_size, fanout, NULL, 0, &edge_size)); }
  void sample_processing_get_co_gpu_omit(VertexId_CUDA *CacheFlag, VertexId_CUDA *dst, VertexId_CUDA *local_column_offset, VertexId_CUDA *global_column_offset,
                                         VertexId_CUDA dst_size, VertexId_CUDA *, VertexId_CUDA, VertexId_CUDA *, VertexId_CUDA *, VertexId_CUDA fanout,
                                         VertexId_CUDA &edge_size) {
    NTS_B200_CHECK(nb_sample_count(ctx, dst, local_column_offset, global_column_offset, dst_size, fanout, CacheFlag, 0xffffffffu, &edge_size)); }
  void sample_processing_get_co_gpu_omit(VertexId_CUDA *CacheFlag, VertexId_CUDA *dst, VertexId_CUDA *local_column_offset, VertexId_CUDA *global_column_offset,
                                         VertexId_CUDA dst_size, VertexId_CUDA *, VertexId_CUDA, VertexId_CUDA *, VertexId_CUDA *, VertexId_CUDA fanout,
                                         VertexId_CUDA &edge_size, VertexId_CUDA super_batch_id) {
    NTS_B200_CHECK(nb_sample_count(ctx, dst, local_column_offset, global_column_offset, dst_size, fanout, CacheFlag, super_batch_id, &edge_size));
    total_sample_num += dst_size; }
  void sample_processing_update_ri_gpu(VertexId_CUDA *r_i, VertexId_CUDA *src_index, VertexId_CUDA edge_size, VertexId_CUDA) {
    NTS_B200_CHECK(nb_sample_update_ri(ctx, r_i, src_index, edge_size)); }
  void sample_processing_traverse_gpu(VertexId_CUDA *destination, VertexId_CUDA *c_o, VertexId_CUDA *r_i, VertexId_CUDA *global_c_o, VertexId_CUDA *global_r_i,
                                      VertexId_CUDA *src_index, VertexId_CUDA vtx_size, VertexId_CUDA edge_size, VertexId_CUDA src_index_size,
                                      VertexId_CUDA *src, VertexId_CUDA *src_count, VertexId_CUDA layer, VertexId_CUDA max_sample_num, bool add_dst_to_src = false) {
    NTS_B200_CHECK(nb_sample_traverse(ctx, destination, c_o, r_i, global_c_o, global_r_i, src_index, vtx_size, edge_size, src_index_size, src, src_count,
                                      layer, max_sample_num, add_dst_to_src ? 1 : 0, rng_seed, rng_counter++)); }
  void set_dst_local_index(VertexId_CUDA *vtx_index, VertexId_CUDA *dev_destination, size_t destination_size, VertexId_CUDA *dev_dst_to_local) {
    NTS_B200_CHECK(nb_set_dst_local_index(ctx, vtx_index, dev_destination, (uint32_t)destination_size, dev_dst_to_local)); }
  void check_dst_local_index(VertexId_CUDA *, size_t, VertexId_CUDA) {}
  void set_total_local_index(VertexId_CUDA *, size_t, VertexId_CUDA *, VertexId_CUDA *, size_t, VertexId_CUDA *, size_t, VertexId_CUDA *, VertexId_CUDA *, VertexId_CUDA *) {
    NTS_B200_UNSUPPORTED("set_total_local_index (only called from commented-out code in the reference)"); }

  /* -- degrees / weights */
  void ReFreshDegree(VertexId_CUDA *out_degree, VertexId_CUDA *in_degree, VertexId_CUDA vertices) {
    NTS_B200_CHECK(nb_update_degree(ctx, out_degree, in_degree, vertices, 0, NULL, NULL, NULL, NULL, 0)); }
  void UpdateDegree(VertexId_CUDA *out_degree, VertexId_CUDA *in_degree, VertexId_CUDA vertices, VertexId_CUDA *destination, VertexId_CUDA *source,
                    VertexId_CUDA *column_offset, VertexId_CUDA *row_indices) {
    NTS_B200_CHECK(nb_update_degree(ctx, out_degree, in_degree, 0, vertices, destination, source, column_offset, row_indices, 0)); }
  void UpdateDegreeCache(VertexId_CUDA *out_degree, VertexId_CUDA *in_degree, VertexId_CUDA vertices, VertexId_CUDA *destination, VertexId_CUDA *source,
                         VertexId_CUDA *column_offset, VertexId_CUDA *row_indices, int fanout) {
    NTS_B200_CHECK(nb_update_degree(ctx, out_degree, in_degree, 0, vertices, destination, source, column_offset, row_indices, fanout)); }
  void GetWeight(float *edge_weight, VertexId_CUDA *out_degree, VertexId_CUDA *in_degree, VertexId_CUDA vertices, VertexId_CUDA *destination,
                 VertexId_CUDA *source, VertexId_CUDA *column_offset, VertexId_CUDA *row_indices) {
    NTS_B200_CHECK(nb_edge_weight(ctx, edge_weight, out_degree, in_degree, vertices, destination, source, column_offset, row_indices, 0)); }
  void GetMeanWeight(float *edge_weight, VertexId_CUDA *out_degree, VertexId_CUDA *in_degree, VertexId_CUDA vertices, VertexId_CUDA *destination,
                     VertexId_CUDA *source, VertexId_CUDA *column_offset, VertexId_CUDA *row_indices) {
    NTS_B200_CHECK(nb_edge_weight(ctx, edge_weight, out_degree, in_degree, vertices, destination, source, column_offset, row_indices, 1)); }
  /* the reference copies the in-degrees into BOTH device arrays (cuda/ntsCUDAGraphOP.cu:2067-2068); each array gets its own here */
  void move_degree_to_gpu(VertexId_CUDA *cpu_in_degree, VertexId_CUDA *cpu_out_degree, VertexId_CUDA *gpu_in_degree, VertexId_CUDA *gpu_out_degree, VertexId_CUDA vertexs) {
    NTS_B200_CHECK(nb_memcpy_h2d(ctx, gpu_in_degree, cpu_in_degree, (size_t)vertexs * sizeof(VertexId_CUDA), 0));
    NTS_B200_CHECK(nb_memcpy_h2d(ctx, gpu_out_degree, cpu_out_degree, (size_t)vertexs * sizeof(VertexId_CUDA), 1)); }

  /* -- gathers */
  void zero_copy_feature_move_gpu(float *dev_feature, float *pinned_host_feature, VertexId_CUDA *src_vertex, VertexId_CUDA feature_size, VertexId_CUDA vertex_size) {
    NTS_B200_CHECK(nb_gather_rows(ctx, dev_feature, pinned_host_feature, src_vertex, vertex_size, feature_size, feature_size, feature_size));
    total_transfer_node += vertex_size; }
  /* FastSampler::load_feature_gpu_cache (core/ntsFastSampler.hpp:263-317) splits the bottom layer's sources into a cold and a hot
   * list on the CPU (positions into dev_source, written to mapped pinned arrays) and issues one call per list: row local_idx[i] of
   * dev_feature <- host table row src_vertex[local_idx[i]] / cache row cache_node_hashmap[src_vertex[local_idx[i]]]. Both synchronise
   * (cuda/ntsCUDAGraphOP.cu:1744-1770): the caller refills the lists for the next batch right away. */
  void zero_copy_feature_move_gpu_cache(float *dev_feature, float *host_pinned_feature, VertexId_CUDA *src_vertex, VertexId_CUDA feature_size,
                                        VertexId_CUDA vertex_size, VertexId_CUDA *local_idx) {
    total_transfer_node += vertex_size;
    NTS_B200_CHECK(nb_gather_rows_indexed(ctx, dev_feature, feature_size, host_pinned_feature, feature_size, src_vertex, local_idx, NULL,
                                          vertex_size, feature_size));
    CUDA_DEVICE_SYNCHRONIZE(); }
  void gather_feature_from_gpu_cache(float *dev_feature, float *dev_cache_feature, VertexId_CUDA *src_vertex, VertexId_CUDA feature_size,
                                     VertexId_CUDA vertex_size, VertexId_CUDA *local_idx, VertexId_CUDA *cache_node_hashmap) {
    NTS_B200_CHECK(nb_gather_rows_indexed(ctx, dev_feature, feature_size, dev_cache_feature, feature_size, src_vertex, local_idx,
                                          cache_node_hashmap, vertex_size, feature_size));
    CUDA_DEVICE_SYNCHRONIZE(); }
  /* new: FastSampler::load_feature_gpu_cache in one call (core/ntsFastSampler.hpp:263-317) */
  void gather_feature_cached(float *dev_feature, float *cold_feature, float *dev_cache_feature, VertexId_CUDA *dev_cache_node_hashmap, VertexId_CUDA *src_vertex,
                             VertexId_CUDA feature_size, VertexId_CUDA vertex_size, VertexId_CUDA *dev_hit_count = NULL) {
    NTS_B200_CHECK(nb_gather_rows_cached(ctx, dev_feature, cold_feature, feature_size, dev_cache_feature, feature_size, dev_cache_node_hashmap, src_vertex,
                                         vertex_size, feature_size, feature_size, dev_hit_count)); }
  void global_copy_label_move_gpu(long *dev_label, long *global_dev_label, VertexId_CUDA *dst_vertex, VertexId_CUDA vertex_size) {
    NTS_B200_CHECK(nb_gather_labels(ctx, (int64_t *)dev_label, (const int64_t *)global_dev_label, dst_vertex, vertex_size)); }
  void zero_copy_embedding_move_gpu(float *dev_feature, float *pinned_host_feature, VertexId_CUDA feature_size, VertexId_CUDA vertex_size) {
    cudaMemcpyAsync(dev_feature, pinned_host_feature, (size_t)feature_size * vertex_size * sizeof(float), cudaMemcpyDefault, stream); }

  /* -- hot-vertex row override (GS_SAMPLE_CACHE / *_PC_MULTI forward) */
  void dev_load_share_embedding(float *dev_embedding, float *share_embedding, VertexId_CUDA *dev_cacheflag, VertexId_CUDA *dev_cachelocation,
                                VertexId_CUDA embedding_size, VertexId_CUDA *destination_vertex, VertexId_CUDA vertex_size, VertexId_CUDA super_batch_id) {
    NTS_B200_CHECK(nb_row_override(ctx, dev_embedding, share_embedding, dev_cacheflag, dev_cachelocation, destination_vertex, vertex_size, embedding_size, super_batch_id)); }
  void dev_load_share_embedding_and_feature(float *dev_feature, float *dev_embedding, float *share_feature, float *share_embedding, VertexId_CUDA *dev_cacheflag,
                                            VertexId_CUDA *dev_cachemap, VertexId_CUDA feature_size, VertexId_CUDA embedding_size, VertexId_CUDA *destination_vertex,
                                            VertexId_CUDA vertex_size) {
    NTS_B200_CHECK(nb_row_override2(ctx, dev_feature, dev_embedding, share_feature, share_embedding, dev_cacheflag, dev_cachemap, destination_vertex, vertex_size,
                                    feature_size, embedding_size, 0xffffffffu)); }
  void dev_load_share_embedding_and_feature(float *dev_feature, float *dev_embedding, float *share_feature, float *share_embedding, VertexId_CUDA *dev_cacheflag,
                                            VertexId_CUDA *dev_cachemap, VertexId_CUDA feature_size, VertexId_CUDA embedding_size, VertexId_CUDA *destination_vertex,
                                            VertexId_CUDA vertex_size, VertexId_CUDA super_batch_id) {
    NTS_B200_CHECK(nb_row_override2(ctx, dev_feature, dev_embedding, share_feature, share_embedding, dev_cacheflag, dev_cachemap, destination_vertex, vertex_size,
                                    feature_size, embedding_size, super_batch_id)); }
  void dev_load_share_aggregate(float *dev_feature, float *share_feature, VertexId_CUDA *dev_cacheflag, VertexId_CUDA *dev_cachemap, VertexId_CUDA feature_size,
                                VertexId_CUDA *destination_vertex, VertexId_CUDA vertex_size) {
    NTS_B200_CHECK(nb_row_override(ctx, dev_feature, share_feature, dev_cacheflag, dev_cachemap, destination_vertex, vertex_size, feature_size, 0xffffffffu)); }
  /* older versioned-cache protocol and debug probes: not on any BASELINE config's live path */
  void dev_load_share_embedding(float *, float *, VertexId_CUDA *, VertexId_CUDA *, VertexId_CUDA, VertexId_CUDA *, VertexId_CUDA) { NTS_B200_UNSUPPORTED("dev_load_share_embedding (cacheflag 2|3 protocol)"); }
  void dev_load_share_embedding(float *, float *, VertexId_CUDA *, VertexId_CUDA *, VertexId_CUDA, VertexId_CUDA *, uint8_t *, uint8_t *, VertexId_CUDA) { NTS_B200_UNSUPPORTED("dev_load_share_embedding (mask protocol)"); }
  void dev_update_share_embedding(float *, float *, VertexId_CUDA *, VertexId_CUDA *, VertexId_CUDA, VertexId_CUDA *, VertexId_CUDA) { NTS_B200_UNSUPPORTED("dev_update_share_embedding"); }
  void dev_update_share_embedding_and_feature(float *, float *, float *, float *, VertexId_CUDA *, VertexId_CUDA *, VertexId_CUDA, VertexId_CUDA, VertexId_CUDA *,
                                              VertexId_CUDA *, VertexId_CUDA *, VertexId_CUDA, VertexId_CUDA) { NTS_B200_UNSUPPORTED("dev_update_share_embedding_and_feature"); }
  void dev_Grad_refresh(float *, float *, VertexId_CUDA *, VertexId_CUDA *, VertexId_CUDA, VertexId_CUDA *, VertexId_CUDA) { NTS_B200_UNSUPPORTED("dev_Grad_refresh"); }
  void dev_Grad_accumulate(float *, float *, VertexId_CUDA *, VertexId_CUDA *, VertexId_CUDA, VertexId_CUDA *, VertexId_CUDA) { NTS_B200_UNSUPPORTED("dev_Grad_accumulate"); }
  void dev_get_X_mask(uint8_t *, VertexId_CUDA *, VertexId_CUDA *, VertexId_CUDA) { NTS_B200_UNSUPPORTED("dev_get_X_mask"); }
  void dev_print_avg_weight(VertexId_CUDA *, VertexId_CUDA *, float *, VertexId_CUDA *, VertexId_CUDA *, float *, VertexId_CUDA *, VertexId_CUDA) { NTS_B200_UNSUPPORTED("dev_print_avg_weight"); }

  static void print_cuda_use() {
    size_t free_byte, total_byte;
    if (cudaMemGetInfo(&free_byte, &total_byte) != cudaSuccess) { printf("Error: cudaMemGetInfo fails\n"); exit(1); }
    std::cout << "Now used GPU memory " << ((double)total_byte - (double)free_byte) / 1024.0 / 1024.0 << "  MB\n";
  }
};

#endif /* TEST_HPP */
