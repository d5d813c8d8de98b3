// gather4.cu -- row gather through TMA tensor maps: cp.async.bulk.tensor.2d tile::gather4 (Blackwell: FOUR table rows, named by four
// row coordinates, per instruction) into shared memory, and one tiled tensor store of the 4 x W tile into the (contiguous) output rows.
//
// Same job as k_gather_rows_tma in gather.cu (FastSampler::load_feature_gpu / zero_copy_feature_move_gpu,
// core/ntsFastSampler.hpp:227-261, kernel cuda/ntsCUDATransferKernel.cuh:154-183), which moves ONE row per cp.async.bulk: for wide rows
// (F = 602: 2.4 KB per row) both issue the same bytes per instruction; for narrow rows (F = 100-128: 400-512 bytes) gather4 moves 4x
// the bytes per load and per store. Selected by nb_set_option("gather_variant", 2); measured against the other variants by
// tools/gather_bench.py (profiles/r2b_gather4_ab.txt), which decides whether it is anybody's default.
//
// Tensor maps: table = 2-D {pitch floats, rows} with box {BW, 1} (the gather4 form: the instruction supplies 4 row coordinates),
// output = 2-D {out_pitch, n_rows} with box {BW, 4}; rows wider than 256 floats are split into k equal column boxes.
// One thread owns one shared-memory slot (4 rows): ids -> k gather4 loads on the slot's mbarrier -> wait -> k tile stores -> next group,
// exactly the structure of k_gather_rows_tma. A group's missing rows (n % 4) repeat the last id; the store clips them at the map's edge.
#include <cuda.h>

#include "common.cuh"

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128)
k_gather_rows_g4(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_out,
                 const uint32_t *__restrict__ ids, uint32_t n_rows, uint32_t box_cols, uint32_t boxes, uint32_t slot_bytes, int slots) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t *bars = (uint64_t *)smem;                   // [slots]
  uint8_t *tiles = smem + ((slots * 8 + 127) & ~127);  // [slots][slot_bytes], slot = boxes x (4 rows x box_cols floats)
  const int t = threadIdx.x;
  if (t >= slots) return;
  const uint32_t bar = smem_addr(&bars[t]);
  const uint32_t slot = smem_addr(tiles + (size_t)t * slot_bytes);
  const uint32_t box_bytes = box_cols * 16u;           // 4 rows x box_cols floats
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const unsigned groups = (n_rows + 3) / 4, stride = gridDim.x * slots;
  uint32_t phase = 0;
  for (unsigned g = blockIdx.x * slots + t; g < groups; g += stride) {
    const unsigned r0 = 4 * g, last = n_rows - 1;
    const int i0 = (int)ids[r0], i1 = (int)ids[min(r0 + 1, last)], i2 = (int)ids[min(r0 + 2, last)], i3 = (int)ids[min(r0 + 3, last)];
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous stores out of this slot have finished reading it
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(box_bytes * boxes) : "memory");
    for (uint32_t b = 0; b < boxes; b++)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                   ::"r"(slot + b * box_bytes), "l"(&map_in), "r"(bar), "r"((int)(b * box_cols)), "r"(i0), "r"(i1), "r"(i2), "r"(i3) : "memory");
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(bar), "r"(phase) : "memory");
    }
    phase ^= 1;
    for (uint32_t b = 0; b < boxes; b++)
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
                   ::"l"(&map_out), "r"(slot + b * box_bytes), "r"((int)(b * box_cols)), "r"((int)r0) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

}  // namespace

// NB_OK when the gather was enqueued; NB_ERR_UNSUPPORTED when this shape cannot go through tensor maps (the caller falls back)
int nb_gather4_launch(nb_ctx *ctx, float *out, uint64_t out_pitch, const float *table, uint64_t table_pitch, const uint32_t *ids_dev,
                      uint32_t n_rows, uint32_t F) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return NB_ERR_UNSUPPORTED;
  // width moved per row: F when its bytes are a multiple of 16, else the whole padded row when both sides share the pitch
  uint32_t W = F;
  if ((W * 4) % 16) { if (table_pitch == out_pitch && (table_pitch * 4) % 16 == 0) W = (uint32_t)table_pitch; else return NB_ERR_UNSUPPORTED; }
  if (((uintptr_t)table | (uintptr_t)out) % 16 || (table_pitch * 4) % 16 || (out_pitch * 4) % 16 || W > table_pitch || W > out_pitch) return NB_ERR_UNSUPPORTED;
  // column boxes: <= 256 floats each, a multiple of 16 bytes per row; a slot's second and later boxes must also start 128-byte aligned
  uint32_t boxes = 1;
  while (boxes <= 64 && !(W % boxes == 0 && W / boxes <= 256 && ((W / boxes) * 4) % 16 == 0 && (boxes == 1 || ((W / boxes) * 16) % 128 == 0))) boxes++;
  if (boxes > 64) return NB_ERR_UNSUPPORTED;
  const uint32_t box_cols = W / boxes;
  CUtensorMap map_in, map_out;
  {
    cuuint64_t dims[2] = {table_pitch, 0x7fffffffull}, strides[1] = {table_pitch * 4};
    cuuint32_t box[2] = {box_cols, 1}, estr[2] = {1, 1};
    if (enc(&map_in, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)table, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return NB_ERR_UNSUPPORTED;
  }
  {
    cuuint64_t dims[2] = {out_pitch, n_rows}, strides[1] = {out_pitch * 4};
    cuuint32_t box[2] = {box_cols, 4}, estr[2] = {1, 1};
    if (enc(&map_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return NB_ERR_UNSUPPORTED;
  }
  const uint32_t slot_bytes = (boxes * box_cols * 16 + 127) / 128 * 128;
  int slots = (int)((200 * 1024) / (slot_bytes + 8));
  if (slots > 128) slots = 128;
  if (slots < 1) return NB_ERR_UNSUPPORTED;
  const size_t smem = ((size_t)slots * 8 + 127) / 128 * 128 + (size_t)slots * slot_bytes;
  static bool attr = false;
  if (!attr) { NB_CUDA(cudaFuncSetAttribute(k_gather_rows_g4, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); attr = true; }
  const unsigned groups = (n_rows + 3) / 4;
  unsigned grid = (groups + slots - 1) / slots;
  if (grid > (unsigned)ctx->sm_count) grid = (unsigned)ctx->sm_count;
  k_gather_rows_g4<<<grid, 128, smem, ctx->stream>>>(map_in, map_out, ids_dev, n_rows, box_cols, boxes, slot_bytes, slots);
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
}
