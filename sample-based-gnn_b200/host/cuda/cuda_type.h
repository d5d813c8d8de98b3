/* Drop-in for the reference's cuda/cuda_type.h (types and launch constants used by core/). */
#ifndef CUDA_TYPE_H
#define CUDA_TYPE_H
#include <stdint.h>
typedef uint32_t VertexId_CUDA;
const int CUDA_NUM_THREADS = 512;
const int CUDA_NUM_BLOCKS = 128;
const int WARP_SIZE = 32;
const int CUDA_NUM_THREADS_SOFTMAX = 32;
const int CUDA_NUM_BLOCKS_SOFTMAX = 512;
#endif /* CUDA_TYPE_H */
