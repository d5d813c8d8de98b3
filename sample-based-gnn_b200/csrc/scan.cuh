// scan.cuh -- device-resident batch state and the single-pass scan shared by the sampler (sample.cu) and the hotness
// pre-sampler (hotness.cu).
#pragma once
#include "common.cuh"

struct nb_graph {
  nb_ctx *ctx;
  uint32_t V;
  uint64_t E;
  uint32_t *col_off, *row_idx, *in_deg, *out_deg;
  uint32_t max_in_degree;
};

struct LayerMeta {  // device resident, one per layer (+1 sentinel)
  uint32_t n_dst, n_edges, n_src, err;
  uint32_t long_rows, pad0, pad1, pad2;
};


// Everything that changes from batch to batch lives in device memory so that the kernel
// arguments are constant and the whole batch can be replayed as one CUDA graph.
struct BatchParams {
  uint64_t rng_seed, rng_offset;
  const uint32_t *omit;  // device, [|V|] or NULL
  uint32_t n_seeds, weight_type, omit_value, replay;
  uint32_t epoch, pad[3];
};

struct ScanWs {
  unsigned long long *tile_state;  // one array per scan launch of a batch; entries are tagged with the batch epoch
  const BatchParams *params;
  unsigned *ticket;                // [2], zero between launches: next tile to hand out, blocks that have left
};
// every caller sizes its tile_state array one entry larger than the largest tile count; that spare entry holds the ticket words
static inline ScanWs nb_scan_ws(unsigned long long *tile_state, size_t entries, const BatchParams *params) {
  return ScanWs{tile_state, params, reinterpret_cast<unsigned *>(tile_state + entries - 1)};
}


// ---------------------------------------------------------------------------------------------
// Single-pass exclusive scan (decoupled look-back, Merrill & Garland) over n items, n read from
// device memory through Op. Tiles are handed out by an atomic ticket, so the block that owns a tile's
// predecessor is always running or finished whatever the grid size and whoever else occupies the SMs
// (a static blockIdx -> tile map can deadlock when the grid is not fully resident); the last block to
// leave zeroes the ticket for the next launch (graph replays included).
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <class Op>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan(Op op, ScanWs ws) {
  __shared__ unsigned s_prefix, s_warp[SCAN_THREADS / 32];
  __shared__ unsigned s_items[SCAN_TILE + SCAN_TILE / 32];
  const unsigned n = op.n();
  const unsigned ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (n == 0 && blockIdx.x == 0 && threadIdx.x == 0) op.total(0);
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // tile state word: [63:34] batch epoch, [33:32] 1 = aggregate, 2 = inclusive prefix, [31:0] value.
  const unsigned long long tag = ((unsigned long long)(ws.params->epoch & 0x3fffffffu)) << 34;
  __shared__ unsigned s_tile;
  while (true) {
    if (threadIdx.x == 0) s_tile = atomicAdd(&ws.ticket[0], 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    if (tile >= ntiles) break;
    // items are produced with coalesced (striped) indices and handed to their owner thread (blocked layout) through
    // shared memory; the +i/32 padding keeps both access patterns nearly conflict free
    unsigned v[SCAN_ITEMS], sum = 0;
    const unsigned first = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
      const unsigned i = k * SCAN_THREADS + threadIdx.x, g = tile * SCAN_TILE + i;
      s_items[i + (i >> 5)] = g < n ? op.load(g) : 0u;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
      const unsigned i = threadIdx.x * SCAN_ITEMS + k;
      v[k] = s_items[i + (i >> 5)];
      sum += v[k];
    }
    unsigned incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned t = __shfl_up_sync(FULL_MASK, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned warp_base = 0, block_total = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; w++) {
      unsigned t = s_warp[w];
      if (w < warp) warp_base += t;
      block_total += t;
    }
    if (warp == 0) {
      // warp-parallel decoupled look-back: lane l inspects tile (first_pred - l); the window slides back by 32
      unsigned prefix = 0;
      if (tile == 0) {
        if (lane == 0) atomicExch(&ws.tile_state[0], tag | (2ull << 32) | block_total);
      } else {
        if (lane == 0) atomicExch(&ws.tile_state[tile], tag | (1ull << 32) | block_total);
        int hi = (int)tile - 1;  // nearest predecessor not yet accounted for
        while (true) {
          const int p = hi - (int)lane;
          unsigned long long st = 0;
          bool ready = true;
          if (p >= 0) {
            st = *((volatile unsigned long long *)&ws.tile_state[p]);
            ready = (st >> 34 << 34) == tag;
          }
          if (!__all_sync(FULL_MASK, ready)) continue;  // some predecessor has not published yet: poll again
          const unsigned flag = p >= 0 ? ((unsigned)(st >> 32) & 3u) : 2u;  // "tile -1" acts as an inclusive prefix of 0
          const unsigned incl_mask = __ballot_sync(FULL_MASK, flag == 2u);
          const unsigned val = p >= 0 ? (unsigned)st : 0u;
          const int stop = incl_mask ? __ffs(incl_mask) - 1 : 31;  // nearest lane holding an inclusive prefix
          prefix += warp_sum_u32(lane <= (unsigned)stop ? val : 0u);
          if (incl_mask) break;
          hi -= 32;
        }
        if (lane == 0) atomicExch(&ws.tile_state[tile], tag | (2ull << 32) | (unsigned long long)(prefix + block_total));
      }
      if (lane == 0) s_prefix = prefix;
    }
    __syncthreads();
    unsigned base = s_prefix + warp_base + (incl - sum);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
      if (first + k < n) op.store(first + k, base, v[k]);
      base += v[k];
    }
    if (tile == ntiles - 1 && threadIdx.x == 0) op.total(s_prefix + block_total);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&ws.ticket[1], 1u) == gridDim.x - 1) { ws.ticket[0] = 0u; ws.ticket[1] = 0u; }
  }
}

