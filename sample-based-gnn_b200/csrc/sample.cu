// sample.cu -- per-layer neighbour sampling, dedup/reindex into the mini-batch CSC, CSR build,
// edge weights. One nb_sampler_sample() call runs every layer of a mini-batch on the stream with
// no host round trip: all sizes (E_i, S_i) live in device memory (LayerMeta) and every kernel is
// a grid-stride kernel over a capacity-sized grid that reads its extent from there.
//
// Replaces (reference file:line)
//   FastSampler::sample_gpu_fast[_omit]            core/ntsFastSampler.hpp:648-915
//   SampledSubgraph::gpu_* arenas / per-layer glue core/FullyRepGraph.hpp:91-131, 258-524
//   sample_processing_get_co_gpu[_omit]  (+ its CPU prefix sum round trip)  cuda/ntsCUDAGraphOP.cu:1246-1559
//   sample_processing_traverse_gpu (stage2 sort-based sampler, stage3 |V| scan) :1584-1659
//   sample_processing_update_ri_gpu :1661-1674, set_dst_local_index :1696-1702
//   ReFreshDegree/UpdateDegree, GetWeight/GetMeanWeight :2015-2113
//   sampCSC::csc_to_csr (CPU only in the reference) core/coocsc.hpp:82-111
// Semantics follow the CPU sampler (ntsFastSampler.hpp:962-1140): `source` ascending by global
// id (bitmap scan order), stable CSR, fp32 weights computed as 1/(sqrtf(out)*sqrtf(in)).
//
// Pipeline per layer (kernel : algorithmic bytes; V/E/S = #dst/#edges/#src of the layer)
//   k_scan<CountOp>   : 12V            min(deg,fanout) + single-pass decoupled look-back scan
//   k_sample          : 12V + 8E       warp per dst, Philox4x32-10 rejection sampling, bitmap marks
//   k_scan<BitmapOp>  : 8|V|/32 + 4S   popcount scan over the bitmap words, emits `source` ascending
//   k_relabel         : 8E + 4E        global -> local ids by popcount rank, CSR histogram
//   k_scan<RowOp>     : 8S             row_offset
//   k_csr_fill/k_csr_rows(+long) : 12E + 12E   atomic fill then per-row rank-sort -> stable CSR
//   k_weights         : 12E + 4E
#include <stdlib.h>

#include "common.cuh"
#include "scan.cuh"

struct LayerBuf {
  uint32_t cap_dst, cap_edges, cap_src;
  uint32_t *destination, *column_offset, *sample_ans, *row_indices, *edge_dst, *source;
  uint32_t *row_offset, *row_count, *row_cursor, *column_indices, *csr_tmp, *csr_to_csc, *long_rows;
  uint32_t *dst_local_id, *src_to_dst;
  uint32_t *gather_idx;           // [cap_edges], bottom layer only: global src id | (the batch reads that source >= gather_keep_min_uses times) << 31
  uint32_t *dst_base, *dst_deg;   // [cap_dst] g_col_off[d] and the in-degree of every dst, written by whoever produced `destination`
                                  // (the previous layer's source emission; layer 0: the sampling kernel itself)
  float *ewf, *ewb;
};

#define NB_MAX_LAYERS 8
static int g_sampler_fused = -1;   // "sampler_fused" / NB_SAMPLER_FUSED: 0 (default) = general kernels (look-back scans), 1 = small-shape kernels where a layer fits
void nb_sampler_set_fused(int on) { g_sampler_fused = on; }
static int g_gather_keep_min = 3;   // "gather_keep_min_uses": sources a batch reads at least this often get the evict_last hint bit
void nb_sampler_set_keep_min(int n) { g_gather_keep_min = n < 1 ? 1 : n; }
// "sampler_tail": bit 0 = the sampling kernel's last block does the bitmap's popcount scan, bit 1 = the relabel kernel's last block does
// the next layer's count scan (and, on the small-shape path, the CSR row offsets). Default 0: measured, a single block's scan of 7-25K
// items (10-25 us) loses to the multi-block look-back scan kernel it replaces (profiles/r2_sampler_tail_ab.txt); kept for small layers.
static int g_sampler_tail = 0;
static int g_sampler_csr_branch = 1;   // "sampler_csr_branch": a layer's CSR kernels on a parallel branch of the captured graph
void nb_sampler_set_csr_branch(int v) { g_sampler_csr_branch = v ? 1 : 0; }
void nb_sampler_set_tail(int v) { g_sampler_tail = v & 3; }
static int g_sampler_two_level = -1;   // "sampler_two_level": -1 (default) = by density, 0 = flat dedup bitmap, 1 = two-level (tests)
void nb_sampler_set_two_level(int mode) { g_sampler_two_level = mode; }
struct nb_sampler {
  nb_ctx *ctx;
  nb_graph *g;
  int L;
  int fanout[NB_MAX_LAYERS];
  uint32_t flags, max_batch;
  LayerBuf lay[NB_MAX_LAYERS];
  LayerMeta *meta_dev;   // [L+1]
  cudaEvent_t ev_fork[NB_MAX_LAYERS], ev_join[NB_MAX_LAYERS];   // CSR branch of the captured graph
  LayerMeta *meta_host;  // the sizes of the batch that nb_sampler_wait / a synchronous sample last completed (points into meta_ring)
  LayerMeta *meta_ring;  // pinned [RING][NB_MAX_LAYERS+1]: every batch in flight copies its sizes into its own slot, so a second
                         // asynchronous nb_sampler_sample before nb_sampler_wait cannot tear the sizes the host reads
  int meta_slot;         // slot of the batch enqueued last
  uint32_t *bitmap[2], *word_rank;  // layer i marks bitmap[i & 1]; its relabel pass clears the other one for layer i + 1
  uint32_t *bitmap_l1[2];           // level 1 of the two-level dedup bitmap (NULL: flat bitmap), n_words_l1 words each
  uint32_t n_words_l1;
  uint32_t n_words;
  unsigned long long *tile_states;  // [3 * L][max_tiles]
  BatchParams *params_dev;
  void *arena;
  uint32_t max_tiles;
  // host staging ring for (params, seeds): pinned, guarded by events
  static const int RING = 8;
  uint8_t *stage[RING];
  cudaEvent_t stage_done[RING];
  cudaEvent_t meta_ready[RING];  // recorded after the sizes of that slot's batch reached meta_ring
  int stage_next;
  uint32_t epoch;
  cudaGraphExec_t graph_exec;
  uint64_t graph_kernels;
  bool use_graph;
  int fused;   // small-shape path: 2 kernels per layer (+2 for a CSR) instead of 4 (+4); see k_sample_fused
};

// counts = min(deg, fanout) (fanout -1: deg), 0 for omitted dst; scan -> column_offset; total -> E.
// Reference: sample_processing_get_co_gpu_kernel[_omit] cuda/ntsCUDATransferKernel.cuh:754-822 and the
// CPU count lambda core/ntsFastSampler.hpp:1001-1009.
struct CountOp {
  const uint32_t *g_col_off, *dst;
  const BatchParams *params;
  uint32_t *col_off;
  LayerMeta *meta;       // this layer's; layer 0 takes its dst count from params, deeper layers from the previous bitmap scan
  const LayerMeta *prev; // previous layer's (NULL for layer 0): an arena overflow there empties every later layer
  uint32_t cap_edges;
  int fanout, bottom;    // omit applies to the bottom layer only (ntsFastSampler.hpp:747-763)
  const uint32_t *dst_deg = nullptr;   // [n] degrees of the dst list when the previous layer's relabel left them (layers >= 1): no random hop
  __device__ unsigned n() const {
    if (prev) return prev->err ? 0u : meta->n_dst;
    return params->n_seeds;
  }
  __device__ unsigned load(unsigned i) const {
    uint32_t deg;
    if (dst_deg) deg = dst_deg[i];
    else { const uint32_t d = dst[i]; deg = g_col_off[d + 1] - g_col_off[d]; }
    uint32_t c = (fanout < 0 || deg < (uint32_t)fanout) ? deg : (uint32_t)fanout;
    const uint32_t *omit = bottom ? params->omit : nullptr;
    if (omit) {
      uint32_t f = omit[dst[i]];
      const uint32_t omit_value = params->omit_value;
      if (omit_value == 0xffffffffu ? (f != 0xffffffffu) : (f == omit_value)) c = 0;
    }
    return c;
  }
  __device__ void store(unsigned i, unsigned excl, unsigned) const { col_off[i] = excl; }
  __device__ void total(unsigned t) const {  // runs exactly once per batch and layer: (re)initialises the layer's meta
    const unsigned nd = n();
    col_off[nd] = t;
    meta->n_dst = nd;
    meta->n_edges = t;
    meta->n_src = 0;
    meta->long_rows = 0;
    meta->err = (prev && prev->err) ? prev->err : (t > cap_edges ? 1u : 0u);
  }
};

// popcount scan over the dedup bitmap; emits `source` in ascending global id (the CPU sampler's
// order, core/ntsFastSampler.hpp:1062-1083) and initialises the per-src scratch.
struct BitmapOp {
  const uint32_t *bitmap;
  uint32_t *word_rank;
  LayerMeta *meta, *next_meta;
  uint32_t n_words, cap_src;
  __device__ unsigned n() const { return n_words; }
  __device__ unsigned load(unsigned w) const { return __popc(bitmap[w]); }
  __device__ void store(unsigned w, unsigned excl, unsigned) const { word_rank[w] = excl; }
  __device__ void total(unsigned t) const {
    meta->n_src = t;
    if (t > cap_src) meta->err = 2;
    next_meta->n_dst = t;
  }
};

// Two-level variant: the items are level-1 words; an item's value is the number of marked vertices under it. The store pass leaves
// word_rank[w0] for every TOUCHED level-0 word (the only ones anybody looks up).
struct Bitmap2Op {
  const uint32_t *bm0, *bm1;
  uint32_t *word_rank;
  LayerMeta *meta, *next_meta;
  uint32_t n_words1, cap_src;
  __device__ unsigned n() const { return n_words1; }
  // the 32 level-0 words under a level-1 word are one 128-byte line: fetched as eight independent 16-byte loads (a bit-by-bit walk
  // would serialise up to 32 dependent-latency loads per item)
  __device__ unsigned load(unsigned w1) const {
    const uint32_t bits = bm1[w1];
    if (!bits) return 0u;
    const uint4 *line = reinterpret_cast<const uint4 *>(bm0 + (size_t)w1 * 32u);
    uint4 q[8];
#pragma unroll
    for (int k = 0; k < 8; k++) q[k] = line[k];
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      s += (bits >> (4 * k) & 1u) ? __popc(q[k].x) : 0u;
      s += (bits >> (4 * k + 1) & 1u) ? __popc(q[k].y) : 0u;
      s += (bits >> (4 * k + 2) & 1u) ? __popc(q[k].z) : 0u;
      s += (bits >> (4 * k + 3) & 1u) ? __popc(q[k].w) : 0u;
    }
    return s;
  }
  __device__ void store(unsigned w1, unsigned excl, unsigned) const {
    const uint32_t bits = bm1[w1];
    if (!bits) return;
    const uint4 *line = reinterpret_cast<const uint4 *>(bm0 + (size_t)w1 * 32u);
    uint4 q[8];
#pragma unroll
    for (int k = 0; k < 8; k++) q[k] = line[k];
    unsigned run = excl;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const uint32_t w[4] = {q[k].x, q[k].y, q[k].z, q[k].w};
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (bits >> (4 * k + j) & 1u) { word_rank[(size_t)w1 * 32u + 4 * k + j] = run; run += __popc(w[j]); }
    }
  }
  __device__ void total(unsigned t) const {
    meta->n_src = t;
    if (t > cap_src) meta->err = 2;
    next_meta->n_dst = t;
  }
};

// one thread per vertex bit: source[rank] = v for every marked v, rank = word prefix + popcount below.
// A warp covers one bitmap word, so the writes of a warp are consecutive.
__global__ void __launch_bounds__(256)
k_emit_sources(const uint32_t *__restrict__ bitmap, const uint32_t *__restrict__ word_rank, uint32_t *__restrict__ source,
               uint32_t *__restrict__ row_count, uint32_t *__restrict__ row_cursor, uint32_t *__restrict__ src_to_dst,
               const LayerMeta *meta, uint32_t n_words, uint32_t *__restrict__ src_index_out = nullptr) {
  if (meta->err) return;
  const unsigned lane = lane_id();
  const unsigned warps = (gridDim.x * blockDim.x) >> 5;
  for (unsigned w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_words; w += warps) {
    const uint32_t bits = bitmap[w];
    if (bits & (1u << lane)) {
      const uint32_t k = word_rank[w] + __popc(bits & ((1u << lane) - 1u));
      source[k] = w * 32u + lane;
      if (row_count) { row_count[k] = 0; row_cursor[k] = 0; }
      if (src_to_dst) src_to_dst[k] = 0xffffffffu;
      if (src_index_out) src_index_out[w * 32u + lane] = k;  // legacy |V|-sized global -> local map
    }
  }
}

struct RowOp {
  const uint32_t *row_count;
  uint32_t *row_offset;
  const LayerMeta *meta;
  __device__ unsigned n() const { return meta->err ? 0u : meta->n_src; }
  __device__ unsigned load(unsigned i) const { return row_count[i]; }
  __device__ void store(unsigned i, unsigned excl, unsigned) const { row_offset[i] = excl; }
  __device__ void total(unsigned t) const { row_offset[meta->err ? 0u : meta->n_src] = t; }
};

// ---------------------------------------------------------------------------------------------
// Neighbour selection: one warp per dst.
//   deg <= fanout (or fanout < 0): every in-neighbour in stored order (ntsFastSampler.hpp:1040-1048)
//   otherwise: `fanout` distinct positions, uniform. Lanes draw positions in parallel; a draw that
//   collides with a held value (or with a lower lane's draw of the same round) is redrawn next
//   round. This is sequential rejection sampling with the iid draw sequence ordered (round, lane),
//   i.e. exactly the CPU sampler's law (ntsFastSampler.hpp:1028-1039): a uniform `fanout`-subset.
//   fanout <= 32 keeps the set in registers (__match_any_sync); larger fanouts use a per-warp
//   open-addressing set in shared memory.
// Philox4x32-10 counter = (dst slot, lane + 32*draw block, layer, rng_offset), key = rng_seed.
// mode 1 (replay): sample_ans was supplied; only edge_dst and the bitmap marks are produced.
// exclusive prefix of one value per thread over the block; *total = block sum. s_warp: 33 words of shared memory.
__device__ __forceinline__ unsigned block_excl_scan(unsigned v, unsigned *s_warp, unsigned *total) {
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  unsigned incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned t = __shfl_up_sync(FULL_MASK, incl, o);
    if (lane >= o) incl += t;
  }
  __syncthreads();   // s_warp may still be read from a previous call
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    unsigned w = lane < nwarps ? s_warp[lane] : 0u, wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned t = __shfl_up_sync(FULL_MASK, wi, o);
      if (lane >= o) wi += t;
    }
    s_warp[lane] = wi - w;
    if (lane == 31) s_warp[32] = wi;
  }
  __syncthreads();
  *total = s_warp[32];
  return s_warp[warp] + incl - v;
}

// in place: s_v[0..n) values -> exclusive prefix sums, s_v[n] = total (returned). Every thread of the block calls it.
__device__ __forceinline__ unsigned block_scan_array(unsigned *s_v, unsigned n, unsigned *s_warp) {
  const unsigned ipt = (n + blockDim.x - 1) / blockDim.x;
  const unsigned a = min(n, threadIdx.x * ipt), b = min(n, a + ipt);
  unsigned sum = 0;
  for (unsigned i = a; i < b; i++) sum += s_v[i];
  unsigned total;
  unsigned run = block_excl_scan(sum, s_warp, &total);
  for (unsigned i = a; i < b; i++) { const unsigned c = s_v[i]; s_v[i] = run; run += c; }
  if (threadIdx.x == 0) s_v[n] = total;
  __syncthreads();
  return total;
}


// Executed by ONE whole block: store(i, exclusive prefix of load(0..i)) for i < n, streaming tiles of blockDim.x * 8 items with a
// running carry; returns the total to every thread. The "last block" tails of the sampling / relabel kernels use it to leave the
// prefix sums the NEXT kernel needs, instead of a scan kernel of its own between the two.
template <class Load, class Store>
__device__ __forceinline__ unsigned block_scan_stream(unsigned n, Load load, Store store, unsigned *s_warp) {
  constexpr unsigned IPT = 8;
  unsigned carry = 0;
  for (unsigned base = 0; base < n; base += blockDim.x * IPT) {
    const unsigned i0 = base + threadIdx.x * IPT;
    unsigned v[IPT], sum = 0;
#pragma unroll
    for (unsigned k = 0; k < IPT; k++) { v[k] = i0 + k < n ? load(i0 + k) : 0u; sum += v[k]; }
    unsigned total;
    unsigned run = carry + block_excl_scan(sum, s_warp, &total);
#pragma unroll
    for (unsigned k = 0; k < IPT; k++) {
      if (i0 + k < n) store(i0 + k, run);
      run += v[k];
    }
    carry += total;
  }
  return carry;
}

// true in exactly one block of the grid: the one whose threads all arrive here last. counter must be 0 before the launch and is 0 again
// afterwards (graph replay). Everything the other blocks wrote before the call is visible to the block that gets `true`.
__device__ __forceinline__ bool last_block_arrives(uint32_t *counter) {
  __shared__ unsigned s_is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const bool last = atomicAdd(counter, 1u) == gridDim.x - 1;
    if (last) *counter = 0u;
    s_is_last = last ? 1u : 0u;
  }
  __syncthreads();
  if (s_is_last) __threadfence();
  return s_is_last != 0u;
}

// tail of a sampling kernel (flat bitmap): BitmapOp's popcount scan by the last block -- word_rank, n_src, capacity check
__device__ __noinline__ void bitmap_rank_tail(const uint32_t *bitmap, uint32_t *word_rank, uint32_t n_words, LayerMeta *meta,
                                                 LayerMeta *next_meta, uint32_t cap_src, bool ok, unsigned *s_warp) {
  if (!last_block_arrives(&meta->pad1)) return;
  if (!ok) { if (threadIdx.x == 0) next_meta->n_dst = 0u; return; }
  const unsigned t = block_scan_stream(n_words, [&](unsigned w) { return (unsigned)__popc(__ldcg(bitmap + w)); },
                                       [&](unsigned w, unsigned excl) { word_rank[w] = excl; }, s_warp);
  if (threadIdx.x == 0) {   // BitmapOp::total
    meta->n_src = t;
    if (t > cap_src) meta->err = 2;
    next_meta->n_dst = t;
  }
}

constexpr int SAMPLE_WARPS = 8;

// Dedup bitmap, two levels when the graph is large: level 0 has one bit per vertex, level 1 one bit per level-0 word, set by the
// thread that turns a zero level-0 word non-zero. Everything after sampling (rank scan, source emission, clearing) then walks the
// |V|/1024 level-1 words and the <= S touched level-0 words instead of all |V|/32 words: O(|V|/1024 + S + E) per layer, which is what
// keeps a 111M-vertex graph's batch at the cost of a 233K-vertex graph's. bitmap_l1 == NULL: flat bitmap (small graphs).
__device__ __forceinline__ void mark_vertex(uint32_t *bitmap, uint32_t *bitmap_l1, uint32_t v) {
  const uint32_t old = atomicOr(&bitmap[v >> 5], 1u << (v & 31));
  if (bitmap_l1 && old == 0u) atomicOr(&bitmap_l1[v >> 10], 1u << ((v >> 5) & 31));
}

// GROUP lanes cooperate on one dst (GROUP = 8, 16 or 32 >= fanout for the register path; 32 for the hash path), so a
// fanout-10 layer keeps two dst per warp busy instead of idling 22 lanes.
template <int GROUP>
__global__ void __launch_bounds__(SAMPLE_WARPS * 32)
k_sample(const uint32_t *__restrict__ g_col_off, const uint32_t *__restrict__ g_row_idx, const uint32_t *__restrict__ dst,
         const uint32_t *__restrict__ col_off, uint32_t *__restrict__ sample_ans, uint32_t *__restrict__ edge_dst,
         uint32_t *__restrict__ bitmap, LayerMeta *meta, int fanout, const BatchParams *params, uint32_t layer,
         int merge, int hash_slots, uint32_t *__restrict__ row_count, uint32_t *__restrict__ row_cursor,
         uint32_t *__restrict__ src_to_dst, uint32_t cap_src, uint32_t *__restrict__ bitmap_l1 = nullptr,
         const uint32_t *__restrict__ dst_base = nullptr, const uint32_t *__restrict__ dst_deg = nullptr,
         uint32_t *__restrict__ tail_word_rank = nullptr, uint32_t tail_words = 0) {
  extern __shared__ uint32_t s_hash[];
  __shared__ unsigned s_warp[33];
  const bool ok = meta->err == 0;   // left by the kernel that produced this layer's column offsets: the same for every block
  if (!ok && !tail_word_rank) return;
  if (ok && row_count) {  // per-src scratch of this layer: S <= E (+V when dst are merged into src)
    const unsigned bound = min(cap_src, meta->n_edges + (merge ? meta->n_dst : 0u));
    for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < bound; k += gridDim.x * blockDim.x) {
      row_count[k] = 0;
      row_cursor[k] = 0;
      if (src_to_dst) src_to_dst[k] = 0xffffffffu;
    }
  }
  const uint64_t key = params->rng_seed ^ (params->rng_offset >> 32 << 32);
  const uint32_t rng_offset = (uint32_t)params->rng_offset;
  const int replay = params->replay;
  const unsigned n_dst = ok ? meta->n_dst : 0u;
  const unsigned lane = lane_id();
  constexpr unsigned GPW = 32 / GROUP;                       // groups per warp
  const unsigned gl = lane % GROUP, gid = lane / GROUP;      // lane in group, group in warp
  const unsigned gmask = GROUP == 32 ? FULL_MASK : (((1u << GROUP) - 1u) << (gid * GROUP));
  const unsigned groups = gridDim.x * SAMPLE_WARPS * GPW;
  uint32_t *my_hash = s_hash + (threadIdx.x >> 5) * hash_slots;
  const Philox rng(key);
  for (unsigned j = (blockIdx.x * SAMPLE_WARPS + (threadIdx.x >> 5)) * GPW + gid; j < n_dst; j += groups) {
    uint32_t base, deg;
    if (dst_base) { base = dst_base[j]; deg = dst_deg[j]; }   // left by the relabel kernel that emitted this dst list: no dependent hop
    else { const uint32_t d = dst[j]; base = g_col_off[d]; deg = g_col_off[d + 1] - base; }
    const uint32_t off = col_off[j];
    const uint32_t num = col_off[j + 1] - off;
    if (merge && gl == 0) mark_vertex(bitmap, bitmap_l1, dst[j]);
    if (num == 0) continue;
    if (replay) {
      for (uint32_t t = gl; t < num; t += GROUP) {
        uint32_t v = sample_ans[off + t];
        edge_dst[off + t] = j;
        mark_vertex(bitmap, bitmap_l1, v);
      }
    } else if (num == deg) {  // take all, stored order
      for (uint32_t t = gl; t < num; t += GROUP) {
        uint32_t v = g_row_idx[base + t];
        sample_ans[off + t] = v;
        edge_dst[off + t] = j;
        mark_vertex(bitmap, bitmap_l1, v);
      }
    } else if (num <= GROUP) {
      const bool holder = gl < num;
      const unsigned holders = __ballot_sync(gmask, holder) & gmask;
      bool need = holder;
      uint32_t pos = 0xffffffffu;
      uint4 r = make_uint4(0, 0, 0, 0);
      for (unsigned round = 0;; round++) {
        if (need) {
          if ((round & 3) == 0) r = rng(j, gl + 32u * (round >> 2), layer, rng_offset);
          uint32_t x = (round & 3) == 0 ? r.x : (round & 3) == 1 ? r.y : (round & 3) == 2 ? r.z : r.w;
          pos = __umulhi(x, deg);
        }
        bool keep = true;
        if (holder) {
          unsigned grp = __match_any_sync(holders, pos);
          unsigned settled = __ballot_sync(holders, !need);
          keep = !need || ((grp & settled) == 0 && lane == (unsigned)(__ffs(grp) - 1));
        }
        need = !keep;
        if (!__any_sync(gmask, need)) break;
      }
      if (holder) {
        uint32_t v = g_row_idx[base + pos];
        sample_ans[off + gl] = v;
        edge_dst[off + gl] = j;
        mark_vertex(bitmap, bitmap_l1, v);
      }
    } else if (GROUP == 32) {  // fanout > 32: shared-memory set, 32 draws per round
      for (int t = lane; t < hash_slots; t += 32) my_hash[t] = 0xffffffffu;
      __syncwarp();
      uint32_t have = 0;
      for (unsigned round = 0; have < num; round++) {
        const uint32_t want = num - have;
        const bool active = lane < want;
        bool won = false;
        uint32_t pos = 0;
        if (active) {
          uint4 r = rng(j, lane + 32u * round, layer, rng_offset);
          pos = __umulhi(r.x, deg);
          uint32_t h = (pos * 2654435761u) & (hash_slots - 1);
          while (true) {
            uint32_t old = atomicCAS(&my_hash[h], 0xffffffffu, pos);
            if (old == 0xffffffffu) { won = true; break; }
            if (old == pos) break;
            h = (h + 1) & (hash_slots - 1);
          }
        }
        unsigned wins = __ballot_sync(FULL_MASK, won);
        if (won) {
          uint32_t slot = have + __popc(wins & ((1u << lane) - 1));
          uint32_t v = g_row_idx[base + pos];
          sample_ans[off + slot] = v;
          edge_dst[off + slot] = j;
          mark_vertex(bitmap, bitmap_l1, v);
        }
        have += __popc(wins);
        __syncwarp();
      }
    }
  }
  if (tail_word_rank) bitmap_rank_tail(bitmap, tail_word_rank, tail_words, meta, meta + 1, cap_src, ok, s_warp);
}

static void launch_sample(cudaStream_t st, unsigned cap_dst, int fanout, const uint32_t *g_col_off, const uint32_t *g_row_idx,
                          const uint32_t *dst, const uint32_t *col_off, uint32_t *sample_ans, uint32_t *edge_dst, uint32_t *bitmap,
                          LayerMeta *meta, const BatchParams *params, uint32_t layer, int merge, uint32_t *row_count = nullptr,
                          uint32_t *row_cursor = nullptr, uint32_t *src_to_dst = nullptr, uint32_t cap_src = 0, uint32_t *bitmap_l1 = nullptr,
                          const uint32_t *dst_base = nullptr, const uint32_t *dst_deg = nullptr, unsigned bps = 8,
                          uint32_t *tail_word_rank = nullptr, uint32_t tail_words = 0) {
  uint32_t hs = 1; while (fanout > 32 && hs < 2u * (uint32_t)fanout) hs <<= 1;
  const int hash_slots = fanout > 32 ? (int)hs : 0;
  const int group = (fanout < 0 || fanout > 16) ? 32 : (fanout > 8 ? 16 : 8);
  const unsigned per_block = SAMPLE_WARPS * (32 / group);
  unsigned grid = nb_grid(cap_dst, per_block, bps);
  if (row_count && grid < NB_SM_COUNT) grid = NB_SM_COUNT;  // enough threads for the scratch clear
  const size_t smem = (size_t)hash_slots * SAMPLE_WARPS * 4;
  if (group == 32) k_sample<32><<<grid, SAMPLE_WARPS * 32, smem, st>>>(g_col_off, g_row_idx, dst, col_off, sample_ans, edge_dst, bitmap, meta, fanout, params, layer, merge, hash_slots, row_count, row_cursor, src_to_dst, cap_src, bitmap_l1, dst_base, dst_deg, tail_word_rank, tail_words);
  else if (group == 16) k_sample<16><<<grid, SAMPLE_WARPS * 32, smem, st>>>(g_col_off, g_row_idx, dst, col_off, sample_ans, edge_dst, bitmap, meta, fanout, params, layer, merge, hash_slots, row_count, row_cursor, src_to_dst, cap_src, bitmap_l1, dst_base, dst_deg, tail_word_rank, tail_words);
  else k_sample<8><<<grid, SAMPLE_WARPS * 32, smem, st>>>(g_col_off, g_row_idx, dst, col_off, sample_ans, edge_dst, bitmap, meta, fanout, params, layer, merge, hash_slots, row_count, row_cursor, src_to_dst, cap_src, bitmap_l1, dst_base, dst_deg, tail_word_rank, tail_words);
}

// global -> local ids: rank(v) = word_rank[v/32] + popc(bitmap[v/32] below bit v%32); CSR histogram.
// Reference: sample_processing_update_ri_gpu_kernel cuda/ntsCUDATransferKernel.cuh:1136-1150,
// sample_set_dst_local :1189-1196; CPU :1085-1099.
struct NextCount {            // k_relabel's tail: the next layer's CountOp (all optional)
  uint32_t *col_off = nullptr;  // [n_src + 1] column offsets of the next layer
  LayerMeta *next_meta = nullptr;
  uint32_t cap_edges = 0;
  int fanout = 0, bottom = 0;
};
__device__ __forceinline__ LayerMeta *meta_mut(const LayerMeta *m) { return const_cast<LayerMeta *>(m); }

__device__ __forceinline__ float edge_weight_fn(uint32_t od, uint32_t id, uint32_t col_len, int weight_type) {
  float w = __fdiv_rn(1.0f, __fmul_rn(__fsqrt_rn((float)od), __fsqrt_rn((float)id)));
  if (weight_type == NB_WEIGHT_MEAN) w = __fdiv_rn(w, (float)id);
  else if (weight_type == NB_WEIGHT_MEAN_SAMPLED) w = __fdiv_rn(w, (float)col_len);
  return w;
}

// fuse_weights: edge weights from the graph's degree arrays are computed in the same pass (UP_DEGREE needs the
// finished histogram and runs k_weights afterwards). Weights: get_weight / get_mean_weight
// cuda/ntsCUDATransferKernel.cuh:294-342, CPU nts_norm_degree core/ntsBaseOp.hpp:652-657. (float)sqrt((double)u32) ==
// sqrtf((float)u32) for u32 < 2^24 (sqrt double rounding is innocuous at 53 >= 2*24+2 bits): bit-identical to the CPU.
__global__ void __launch_bounds__(256)
k_relabel(const uint32_t *__restrict__ sample_ans, uint32_t *__restrict__ row_indices, const uint32_t *__restrict__ bitmap,
          const uint32_t *__restrict__ word_rank, uint32_t *__restrict__ row_count, const uint32_t *__restrict__ dst,
          uint32_t *__restrict__ dst_local_id, uint32_t *__restrict__ src_to_dst, const LayerMeta *meta, int histogram,
          int fuse_weights, float *__restrict__ ewf, const uint32_t *__restrict__ edge_dst, const uint32_t *__restrict__ col_off,
          const uint32_t *__restrict__ in_deg, const uint32_t *__restrict__ out_deg, const BatchParams *params,
          uint32_t *__restrict__ source, uint32_t n_words, uint32_t *__restrict__ other_bitmap,
          const uint32_t *__restrict__ g_col_off = nullptr, uint32_t *__restrict__ next_base = nullptr,
          uint32_t *__restrict__ next_deg = nullptr, const uint32_t *__restrict__ bitmap_l1 = nullptr,
          uint32_t *__restrict__ other_bitmap_l1 = nullptr, uint32_t n_words_l1 = 0, NextCount tail = NextCount{}) {
  __shared__ unsigned s_warp[33];
  const unsigned stride = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
  // the next layer marks the other bitmap: clear it here, off the critical path (no memset node per layer)
  if (other_bitmap) {
    if (bitmap_l1) {   // two levels: only the touched level-0 words are cleared, found through level 1
      const unsigned lane = lane_id();
      for (unsigned w1 = tid >> 5; w1 < n_words_l1; w1 += stride >> 5) {
        const uint32_t bits1 = other_bitmap_l1[w1];
        if (bits1 >> lane & 1u) other_bitmap[w1 * 32u + lane] = 0u;
        __syncwarp();
        if (lane == 0 && bits1) other_bitmap_l1[w1] = 0u;
      }
    } else {
      for (unsigned w = tid; w <= n_words; w += stride) other_bitmap[w] = 0u;
    }
  }
  const unsigned err_in = meta->err;   // left by the kernels before this one: the same for every block
  if (err_in && !tail.col_off) return;
  const unsigned E = err_in ? 0u : meta->n_edges, nd = err_in ? 0u : meta->n_dst;
  const int weight_type = params->weight_type;
  if (err_in) {
  } else if (bitmap_l1) {   // `source` ascending: warp per level-1 word, lane per touched level-0 word, a short serial walk over its bits
    const unsigned lane = lane_id();
    for (unsigned w1 = tid >> 5; w1 < n_words_l1; w1 += stride >> 5) {
      if (bitmap_l1[w1] >> lane & 1u) {
        const uint32_t w0 = w1 * 32u + lane;
        uint32_t bits = bitmap[w0], k = word_rank[w0];
        while (bits) {
          const uint32_t v = w0 * 32u + (__ffs(bits) - 1);
          bits &= bits - 1;
          source[k] = v;
          if (next_base) { const uint32_t b = g_col_off[v]; next_base[k] = b; next_deg[k] = g_col_off[v + 1] - b; }
          k++;
        }
      }
    }
  } else
  // `source` in ascending global id: one thread per vertex bit, a warp covers one bitmap word (consecutive writes)
  {
    const unsigned lane = lane_id();
    for (unsigned w = tid >> 5; w < n_words; w += stride >> 5) {
      const uint32_t bits = bitmap[w];
      if (bits & (1u << lane)) {
        const uint32_t k = word_rank[w] + __popc(bits & ((1u << lane) - 1u)), v = w * 32u + lane;
        source[k] = v;
        if (next_base) {   // the next layer's dst list is this `source`: its sampler finds base / degree without touching g_col_off
          const uint32_t b = g_col_off[v];
          next_base[k] = b;
          next_deg[k] = g_col_off[v + 1] - b;
        }
      }
    }
  }
  for (unsigned e = tid; e < E; e += stride) {
    uint32_t v = sample_ans[e];
    uint32_t local = word_rank[v >> 5] + __popc(bitmap[v >> 5] & ((1u << (v & 31)) - 1u));
    row_indices[e] = local;
    if (histogram) atomicAdd(&row_count[local], 1u);
    if (fuse_weights && weight_type != NB_WEIGHT_NONE) {
      const uint32_t j = edge_dst[e];
      ewf[e] = edge_weight_fn(out_deg[v], in_deg[dst[j]], col_off[j + 1] - col_off[j], weight_type);
    }
  }
  if (dst_local_id)
    for (unsigned j = tid; j < nd; j += stride) {
      uint32_t v = dst[j];
      uint32_t local = word_rank[v >> 5] + __popc(bitmap[v >> 5] & ((1u << (v & 31)) - 1u));
      dst_local_id[j] = local;
      src_to_dst[local] = j;
    }
  // tail: the block that finishes last leaves the NEXT layer's column offsets and meta (CountOp's scan; that layer's dst list is this
  // `source`, its degrees are in next_deg), so the next sampling kernel starts without a scan kernel in between
  if (tail.col_off && last_block_arrives(&meta_mut(meta)->pad0)) {
    LayerMeta *nm = tail.next_meta;
    if (err_in) {
      if (threadIdx.x == 0) { nm->n_dst = 0; nm->n_edges = 0; nm->n_src = 0; nm->long_rows = 0; nm->err = err_in; tail.col_off[0] = 0u; }
      return;
    }
    const unsigned S = meta->n_src;
    const uint32_t *omit = tail.bottom ? params->omit : nullptr;
    const uint32_t omit_value = params->omit_value;
    const int fanout = tail.fanout;
    const unsigned t = block_scan_stream(S, [&](unsigned k) {
      const uint32_t deg = __ldcg(next_deg + k);
      uint32_t c = (fanout < 0 || deg < (uint32_t)fanout) ? deg : (uint32_t)fanout;
      if (omit) {
        const uint32_t f = omit[__ldcg(source + k)];
        if (omit_value == 0xffffffffu ? (f != 0xffffffffu) : (f == omit_value)) c = 0;
      }
      return c; }, [&](unsigned k, unsigned excl) { tail.col_off[k] = excl; }, s_warp);
    if (threadIdx.x == 0) {   // CountOp::total
      tail.col_off[S] = t;
      nm->n_dst = S; nm->n_edges = t; nm->n_src = 0; nm->long_rows = 0;
      nm->err = t > tail.cap_edges ? 1u : 0u;
    }
  }
}

// clears a two-level bitmap through its level 1 (odd layer counts: the last layer leaves bitmap[0] marked for the next batch's layer 0)
__global__ void __launch_bounds__(256)
k_clear_two_level(uint32_t *__restrict__ bitmap, uint32_t *__restrict__ bitmap_l1, uint32_t n_words_l1) {
  const unsigned lane = lane_id();
  for (unsigned w1 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w1 < n_words_l1; w1 += (gridDim.x * blockDim.x) >> 5) {
    const uint32_t bits1 = bitmap_l1[w1];
    if (bits1 >> lane & 1u) bitmap[w1 * 32u + lane] = 0u;
    __syncwarp();
    if (lane == 0 && bits1) bitmap_l1[w1] = 0u;
  }
}

// UP_DEGREE weights: sampled degrees (out = CSR row length, in = CSC column length), core/FullyRepGraph.hpp:189-207.
__global__ void __launch_bounds__(256)
k_weights_sampled(float *__restrict__ ewf, const uint32_t *__restrict__ row_indices, const uint32_t *__restrict__ edge_dst,
                  const uint32_t *__restrict__ col_off, const uint32_t *__restrict__ row_count, const LayerMeta *meta,
                  const BatchParams *params) {
  if (meta->err) return;
  const int weight_type = params->weight_type;
  if (weight_type == NB_WEIGHT_NONE) return;
  const unsigned E = meta->n_edges;
  for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    const uint32_t j = edge_dst[e];
    const uint32_t col_len = col_off[j + 1] - col_off[j];
    ewf[e] = edge_weight_fn(row_count[row_indices[e]], col_len, col_len, weight_type);
  }
}

// CSR build, step 1: drop every edge into its row in arrival order.
__global__ void __launch_bounds__(256)
k_csr_fill(const uint32_t *__restrict__ row_indices, const uint32_t *__restrict__ row_offset, uint32_t *__restrict__ row_cursor,
           uint32_t *__restrict__ csr_tmp, const LayerMeta *meta) {
  if (meta->err) return;
  const unsigned E = meta->n_edges;
  for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    uint32_t s = row_indices[e];
    csr_tmp[row_offset[s] + atomicAdd(&row_cursor[s], 1u)] = e;
  }
}

// step 2: order each row by CSC position (== ascending local dst, then position inside the column:
// exactly sampCSC::csc_to_csr's stable fill, core/coocsc.hpp:98-105) by rank counting, and emit
// column_indices / csr_to_csc / e_w_b. Short rows: one thread each; long rows are queued for a warp.
constexpr uint32_t CSR_SHORT = 16;
__device__ __forceinline__ void csr_emit(uint32_t pos, uint32_t e, uint32_t *column_indices, uint32_t *csr_to_csc,
                                         const uint32_t *edge_dst, float *ewb, const float *ewf) {
  csr_to_csc[pos] = e;
  column_indices[pos] = edge_dst[e];
  if (ewb) ewb[pos] = ewf[e];
}
__global__ void __launch_bounds__(256)
k_csr_rows(const uint32_t *__restrict__ row_offset, const uint32_t *__restrict__ csr_tmp, uint32_t *__restrict__ column_indices,
           uint32_t *__restrict__ csr_to_csc, const uint32_t *__restrict__ edge_dst, float *__restrict__ ewb,
           const float *__restrict__ ewf, uint32_t *__restrict__ long_rows, LayerMeta *meta, const BatchParams *params,
           int inline_long = 0) {
  if (meta->err) return;
  if (params->weight_type == NB_WEIGHT_NONE) ewb = nullptr;
  const unsigned S = meta->n_src;
  const unsigned lane = lane_id();
  const unsigned S_round = (S + 31u) & ~31u;   // whole warps stay together for the cooperative long-row pass
  for (unsigned s = blockIdx.x * blockDim.x + threadIdx.x; s < S_round; s += gridDim.x * blockDim.x) {
    uint32_t a = 0, n = 0;
    if (s < S) { a = row_offset[s]; n = row_offset[s + 1] - a; }
    const bool is_long = n > CSR_SHORT;
    if (is_long && !inline_long) long_rows[atomicAdd(&meta->long_rows, 1u)] = s;  // hub source: queued for k_csr_long_rows
    if (!is_long && n) {
      uint32_t ev[CSR_SHORT];
#pragma unroll
      for (uint32_t i = 0; i < CSR_SHORT; i++) ev[i] = i < n ? csr_tmp[a + i] : 0xffffffffu;
#pragma unroll
      for (uint32_t i = 0; i < CSR_SHORT; i++) {
        if (i < n) {
          uint32_t rank = 0;
#pragma unroll
          for (uint32_t k = 0; k < CSR_SHORT; k++) rank += (ev[k] < ev[i]);
          csr_emit(a + rank, ev[i], column_indices, csr_to_csc, edge_dst, ewb, ewf);
        }
      }
    }
    if (inline_long) {   // a hub source is ordered by the whole warp (rank counting, O(n^2/32)), not by one thread
      unsigned pending = __ballot_sync(FULL_MASK, is_long);
      while (pending) {
        const int owner = __ffs(pending) - 1;
        pending &= pending - 1;
        const uint32_t ra = __shfl_sync(FULL_MASK, a, owner), rn = __shfl_sync(FULL_MASK, n, owner);
        for (uint32_t i = lane; i < rn; i += 32) {
          const uint32_t e = csr_tmp[ra + i];
          uint32_t rank = 0;
          for (uint32_t k = 0; k < rn; k++) rank += (__ldg(&csr_tmp[ra + k]) < e);
          csr_emit(ra + rank, e, column_indices, csr_to_csc, edge_dst, ewb, ewf);
        }
      }
    }
  }
}
// one warp per queued long row, rank counting over the row (O(n^2/32), rows of a few hundred entries at most in practice)
__global__ void __launch_bounds__(256)
k_csr_long_rows(const uint32_t *__restrict__ row_offset, const uint32_t *__restrict__ csr_tmp, uint32_t *__restrict__ column_indices,
                uint32_t *__restrict__ csr_to_csc, const uint32_t *__restrict__ edge_dst, float *__restrict__ ewb,
                const float *__restrict__ ewf, const uint32_t *__restrict__ long_rows, const LayerMeta *meta,
                const BatchParams *params) {
  if (meta->err) return;
  if (params->weight_type == NB_WEIGHT_NONE) ewb = nullptr;
  const unsigned n_long = meta->long_rows;
  const unsigned lane = lane_id();
  for (unsigned w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_long; w += (gridDim.x * blockDim.x) >> 5) {
    const uint32_t s = long_rows[w];
    const uint32_t a = row_offset[s], n = row_offset[s + 1] - a;
    for (uint32_t i = lane; i < n; i += 32) {
      const uint32_t e = csr_tmp[a + i];
      uint32_t rank = 0;
      for (uint32_t k = 0; k < n; k++) rank += (__ldg(&csr_tmp[a + k]) < e);
      csr_emit(a + rank, e, column_indices, csr_to_csc, edge_dst, ewb, ewf);
    }
  }
}

// =============================================================================================
// Small-shape path ("sampler_fused" = 1; NOT the default). When the per-layer arrays fit in shared memory every block recomputes the
// layer's prefix sums for itself (a few tens of KB from L2) instead of waiting for a separate scan kernel:
//   k_sample_fused   = count + exclusive scan (in smem, per block) + neighbour selection + bitmap marks
//   k_relabel_fused  = popcount scan of the dedup bitmap (in smem, per block) + `source` emission + relabel + histogram +
//                      weights (+ the next layer's per-dst adjacency base / degree, so its sampler skips one dependent load)
//   k_csr_fill_fused = row_offset scan (in smem, per block) + stable-fill step 1;  k_csr_rows handles long rows itself
// i.e. 7 launches instead of 12 for the benchmark's two layers. Results are bit-identical to the general path (same arithmetic, same
// ordering rules); the tests run every variant. It was the default for half a round. Measured again once batches were sampled on two
// streams beside the aggregation of earlier batches, it lost on both counts: alone (110-135 us per batch vs 80-90: every block redoes a
// 25K-entry scan) and as a neighbour (blocks with 100 KB of shared memory / 512 threads do not fit the slot the aggregation leaves free,
// so they displace its blocks: 0.151-0.167 vs 0.139 ms per step) -- profiles/r2_sweep_sampler_residency.txt.
constexpr int FS_THREADS = 512;   // upper bound of the small-shape kernels' block size; the launch uses g_sampler_block threads
// "sampler_block_threads": 256 (default) or 512. The sampler runs beside the previous batch's aggregation, whose two resident
// blocks per SM leave ~11K registers: a 256-thread block of <= 40 registers fits there, a 512-thread block waits for a retiring one.
static int g_sampler_block = 256;
// "sampler_blocks_per_sm": 0 = every kernel sized for its own work (up to 8 blocks per SM); n > 0 = at most n blocks per SM for every
// kernel of the batch graph: with 1, the sampler lives entirely in the slot the aggregation leaves free and never displaces its blocks
static int g_sampler_bps = 2;
void nb_sampler_set_bps(int v) { g_sampler_bps = v < 0 ? 0 : v; }
static int g_sampler_capture_prio = 1;   // "sampler_capture_priority": capture streams carry the launch stream's priority
void nb_sampler_set_capture_prio(int v) { g_sampler_capture_prio = v; }
static inline unsigned sgrid(uint64_t work, unsigned items, unsigned bps) {
  return nb_grid(work, items, g_sampler_bps > 0 && (unsigned)g_sampler_bps < bps ? (unsigned)g_sampler_bps : bps);
}
void nb_sampler_set_block(int v) { g_sampler_block = v >= 512 ? 512 : 256; }
constexpr size_t FS_SMEM_MAX = 200 * 1024;
constexpr uint32_t FS_TAIL_MAX = 49152;
constexpr uint32_t FS_RANK_TAIL_MAX = 16384, FS_COUNT_TAIL_MAX = 32768;   // longest scans left to a kernel's last block (general path tails)   // per-source arrays up to this long are scanned by the relabel kernel's last block

// One sampling layer's count + scan + neighbour selection. Selection code and RNG counters are those of k_sample: same draws.
template <int GROUP>
__global__ void __launch_bounds__(FS_THREADS)
k_sample_fused(const uint32_t *__restrict__ g_col_off, const uint32_t *__restrict__ g_row_idx, const uint32_t *__restrict__ dst,
               uint32_t *__restrict__ dst_base, uint32_t *__restrict__ dst_deg, int dense, uint32_t *__restrict__ col_off,
               uint32_t *__restrict__ sample_ans, uint32_t *__restrict__ edge_dst, uint32_t *__restrict__ bitmap, LayerMeta *meta,
               const LayerMeta *prev, int fanout, const BatchParams *params, uint32_t layer, int merge, int bottom, int hash_slots,
               uint32_t *__restrict__ row_count, uint32_t *__restrict__ row_cursor, uint32_t *__restrict__ src_to_dst,
               uint32_t cap_src, uint32_t cap_edges, uint32_t cap_dst, uint32_t *__restrict__ bitmap_l1,
               uint32_t *__restrict__ tail_word_rank = nullptr, uint32_t tail_words = 0) {
  extern __shared__ uint32_t s_dyn[];   // [cap_dst + 1] counts -> offsets, then the per-warp hash sets (fanout > 32)
  __shared__ unsigned s_warp[33];
  uint32_t *s_off = s_dyn;
  uint32_t *s_hash = s_dyn + ((cap_dst + 1 + 31) & ~31u);
  const unsigned n_dst = prev ? (prev->err ? 0u : meta->n_dst) : params->n_seeds;
  const uint32_t *omit = bottom ? params->omit : nullptr;
  const uint32_t omit_value = params->omit_value;
  const int replay = params->replay;
  // 1. counts (every block, redundantly: n_dst words from L2), exactly CountOp::load
  for (unsigned i = threadIdx.x; i < n_dst; i += blockDim.x) {
    uint32_t deg;
    if (dense) deg = dst_deg[i];
    else {
      const uint32_t d = dst[i], b = g_col_off[d];
      deg = g_col_off[d + 1] - b;
      if (blockIdx.x == 0) { dst_base[i] = b; dst_deg[i] = deg; }
    }
    uint32_t c = (fanout < 0 || deg < (uint32_t)fanout) ? deg : (uint32_t)fanout;
    if (omit) {
      const uint32_t f = omit[dst[i]];
      if (omit_value == 0xffffffffu ? (f != 0xffffffffu) : (f == omit_value)) c = 0;
    }
    s_off[i] = c;
  }
  __syncthreads();
  // 2. exclusive scan -> column offsets
  const unsigned E = block_scan_array(s_off, n_dst, s_warp);
  const unsigned err = (prev && prev->err) ? prev->err : (E > cap_edges ? 1u : 0u);
  if (blockIdx.x == 0) {
    for (unsigned i = threadIdx.x; i <= n_dst; i += blockDim.x) col_off[i] = s_off[i];
    if (threadIdx.x == 0) { meta->n_dst = n_dst; meta->n_edges = E; meta->n_src = 0; meta->long_rows = 0; meta->err = err; }
  }
  if (err && !tail_word_rank) return;
  if (!err && row_count) {  // per-src scratch of this layer: S <= E (+V when dst are merged into src)
    const unsigned bound = min(cap_src, E + (merge ? n_dst : 0u));
    for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < bound; k += gridDim.x * blockDim.x) {
      row_count[k] = 0;
      row_cursor[k] = 0;
      if (src_to_dst) src_to_dst[k] = 0xffffffffu;
    }
  }
  // 3. neighbour selection (see k_sample for the rule and the RNG counter layout)
  const uint64_t key = params->rng_seed ^ (params->rng_offset >> 32 << 32);
  const uint32_t rng_offset = (uint32_t)params->rng_offset;
  const unsigned lane = lane_id();
  constexpr unsigned GPW = 32 / GROUP;
  const unsigned gl = lane % GROUP, gid = lane / GROUP;
  const unsigned gmask = GROUP == 32 ? FULL_MASK : (((1u << GROUP) - 1u) << (gid * GROUP));
  const unsigned groups = gridDim.x * (blockDim.x / 32) * GPW;
  uint32_t *my_hash = s_hash + (threadIdx.x >> 5) * hash_slots;
  const Philox rng(key);
  for (unsigned j = (blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5)) * GPW + gid; j < (err ? 0u : n_dst); j += groups) {
    const uint32_t off = s_off[j];
    const uint32_t num = s_off[j + 1] - off;
    uint32_t base, deg;
    if (dense) { base = dst_base[j]; deg = dst_deg[j]; }
    else { const uint32_t d = dst[j]; base = g_col_off[d]; deg = g_col_off[d + 1] - base; }
    if (merge && gl == 0) { const uint32_t d = dst[j]; mark_vertex(bitmap, bitmap_l1, d); }
    if (num == 0) continue;
    if (replay) {
      for (uint32_t t = gl; t < num; t += GROUP) {
        uint32_t v = sample_ans[off + t];
        edge_dst[off + t] = j;
        mark_vertex(bitmap, bitmap_l1, v);
      }
    } else if (num == deg) {  // take all, stored order
      for (uint32_t t = gl; t < num; t += GROUP) {
        uint32_t v = g_row_idx[base + t];
        sample_ans[off + t] = v;
        edge_dst[off + t] = j;
        mark_vertex(bitmap, bitmap_l1, v);
      }
    } else if (num <= GROUP) {
      const bool holder = gl < num;
      const unsigned holders = __ballot_sync(gmask, holder) & gmask;
      bool need = holder;
      uint32_t pos = 0xffffffffu;
      uint4 r = make_uint4(0, 0, 0, 0);
      for (unsigned round = 0;; round++) {
        if (need) {
          if ((round & 3) == 0) r = rng(j, gl + 32u * (round >> 2), layer, rng_offset);
          uint32_t x = (round & 3) == 0 ? r.x : (round & 3) == 1 ? r.y : (round & 3) == 2 ? r.z : r.w;
          pos = __umulhi(x, deg);
        }
        bool keep = true;
        if (holder) {
          unsigned grp = __match_any_sync(holders, pos);
          unsigned settled = __ballot_sync(holders, !need);
          keep = !need || ((grp & settled) == 0 && lane == (unsigned)(__ffs(grp) - 1));
        }
        need = !keep;
        if (!__any_sync(gmask, need)) break;
      }
      if (holder) {
        uint32_t v = g_row_idx[base + pos];
        sample_ans[off + gl] = v;
        edge_dst[off + gl] = j;
        mark_vertex(bitmap, bitmap_l1, v);
      }
    } else if (GROUP == 32) {  // fanout > 32: shared-memory set, 32 draws per round
      for (int t = lane; t < hash_slots; t += 32) my_hash[t] = 0xffffffffu;
      __syncwarp();
      uint32_t have = 0;
      for (unsigned round = 0; have < num; round++) {
        const uint32_t want = num - have;
        const bool active = lane < want;
        bool won = false;
        uint32_t pos = 0;
        if (active) {
          uint4 r = rng(j, lane + 32u * round, layer, rng_offset);
          pos = __umulhi(r.x, deg);
          uint32_t h = (pos * 2654435761u) & (hash_slots - 1);
          while (true) {
            uint32_t old = atomicCAS(&my_hash[h], 0xffffffffu, pos);
            if (old == 0xffffffffu) { won = true; break; }
            if (old == pos) break;
            h = (h + 1) & (hash_slots - 1);
          }
        }
        unsigned wins = __ballot_sync(FULL_MASK, won);
        if (won) {
          uint32_t slot = have + __popc(wins & ((1u << lane) - 1));
          uint32_t v = g_row_idx[base + pos];
          sample_ans[off + slot] = v;
          edge_dst[off + slot] = j;
          mark_vertex(bitmap, bitmap_l1, v);
        }
        have += __popc(wins);
        __syncwarp();
      }
    }
  }
  if (tail_word_rank) bitmap_rank_tail(bitmap, tail_word_rank, tail_words, meta, meta + 1, cap_src, err == 0, s_warp);
}

struct RelabelTail {          // what the last block of k_relabel_fused leaves behind (all optional)
  int enabled;
  uint32_t *next_col_off;     // [n_src + 1] column offsets of the NEXT layer (its dst list is this layer's source), or NULL
  uint32_t next_cap_edges;
  int next_fanout, next_bottom;
  uint32_t *row_offset;       // [n_src + 1] CSR row offsets of THIS layer, or NULL
};

// dedup + relabel of one layer: bitmap ranks in shared memory, `source` ascending, local ids, CSR histogram, weights, and the next
// layer's per-dst adjacency base / degree (its dst list IS this `source`). Same results as k_scan<BitmapOp> + k_relabel.
__global__ void __launch_bounds__(FS_THREADS, 3)
k_relabel_fused(const uint32_t *__restrict__ sample_ans, uint32_t *__restrict__ row_indices, const uint32_t *__restrict__ bitmap,
                uint32_t *__restrict__ row_count, const uint32_t *__restrict__ dst, uint32_t *__restrict__ dst_local_id,
                uint32_t *__restrict__ src_to_dst, LayerMeta *meta, LayerMeta *next_meta, int histogram, int fuse_weights,
                float *__restrict__ ewf, const uint32_t *__restrict__ edge_dst, const uint32_t *__restrict__ col_off,
                const uint32_t *__restrict__ in_deg, const uint32_t *__restrict__ out_deg, const BatchParams *params,
                uint32_t *__restrict__ source, uint32_t n_words, uint32_t *__restrict__ other_bitmap, uint32_t cap_src,
                const uint32_t *__restrict__ g_col_off, uint32_t *__restrict__ next_base, uint32_t *__restrict__ next_deg,
                RelabelTail tail) {
  extern __shared__ uint32_t s_dyn[];   // [n_words] bits, [n_words + 1] ranks; the tail reuses it for [n_src + 1] counts
  __shared__ unsigned s_warp[33];
  __shared__ unsigned s_last;
  uint32_t *s_bits = s_dyn, *s_rank = s_dyn + ((n_words + 31) & ~31u);
  const unsigned stride = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (other_bitmap)
    for (unsigned w = tid; w <= n_words; w += stride) other_bitmap[w] = 0u;
  const unsigned err_in = meta->err;      // left by this layer's sampling kernel; the same for every block
  unsigned S = 0;
  if (!err_in) {
    for (unsigned w = threadIdx.x; w < n_words; w += blockDim.x) {
      const uint32_t b = bitmap[w];
      s_bits[w] = b;
      s_rank[w] = __popc(b);
    }
    __syncthreads();
    S = block_scan_array(s_rank, n_words, s_warp);
    if (tid == 0) {
      meta->n_src = S;
      if (S > cap_src) meta->err = 2;
      next_meta->n_dst = S;
    }
  }
  const bool ok = !err_in && S <= cap_src;
  const unsigned E = meta->n_edges, nd = meta->n_dst;
  const int weight_type = params->weight_type;
  if (ok) {
    const unsigned lane = lane_id();
    for (unsigned w = tid >> 5; w < n_words; w += stride >> 5) {
      const uint32_t bits = s_bits[w];
      if (bits & (1u << lane)) {
        const uint32_t k = s_rank[w] + __popc(bits & ((1u << lane) - 1u)), v = w * 32u + lane;
        source[k] = v;
        if (next_base) {   // consecutive lanes hold consecutive vertices: coalesced
          const uint32_t b = g_col_off[v];
          next_base[k] = b;
          next_deg[k] = g_col_off[v + 1] - b;
        }
      }
    }
  }
  if (ok) {
    const bool weights = fuse_weights && weight_type != NB_WEIGHT_NONE;
    constexpr int U = 4;   // edges per thread in flight: the chain sample_ans -> out_deg / edge_dst -> dst -> in_deg is all latency
    for (unsigned e0 = tid; e0 < E; e0 += U * stride) {
      uint32_t v[U], j[U], od[U], cb[U], ce[U], d[U], id[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const unsigned e = e0 + u * stride;
        v[u] = e < E ? sample_ans[e] : 0u;
        j[u] = (weights && e < E) ? edge_dst[e] : 0u;
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        if (weights && e0 + u * stride < E) { od[u] = out_deg[v[u]]; d[u] = dst[j[u]]; cb[u] = col_off[j[u]]; ce[u] = col_off[j[u] + 1]; }
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        if (weights && e0 + u * stride < E) id[u] = in_deg[d[u]];
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const unsigned e = e0 + u * stride;
        if (e < E) {
          const uint32_t local = s_rank[v[u] >> 5] + __popc(s_bits[v[u] >> 5] & ((1u << (v[u] & 31)) - 1u));
          row_indices[e] = local;
          if (histogram) atomicAdd(&row_count[local], 1u);
          if (weights) ewf[e] = edge_weight_fn(od[u], id[u], ce[u] - cb[u], weight_type);
        }
      }
    }
  }
  if (ok && dst_local_id)
    for (unsigned j = tid; j < nd; j += stride) {
      const uint32_t v = dst[j];
      const uint32_t local = s_rank[v >> 5] + __popc(s_bits[v >> 5] & ((1u << (v & 31)) - 1u));
      dst_local_id[j] = local;
      src_to_dst[local] = j;
    }
  if (!tail.enabled) return;
  // ---- tail: the block that finishes last turns the finished per-source arrays into prefix sums, so that neither the next layer's
  // sampling kernel nor this layer's CSR fill needs a scan of its own (every one of their blocks used to redo it in shared memory)
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&meta->pad0, 1u) == gridDim.x - 1 ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x == 0) meta->pad0 = 0u;   // re-armed for the next batch (graph replay)
  if (tail.next_col_off) {   // the next layer's CountOp: its dst list is this `source`, the degrees are in next_deg
    unsigned total = 0;
    if (ok) {
      const uint32_t *omit = tail.next_bottom ? params->omit : nullptr;
      const uint32_t omit_value = params->omit_value;
      for (unsigned k = threadIdx.x; k < S; k += blockDim.x) {
        const uint32_t deg = __ldcg(next_deg + k);
        uint32_t c = (tail.next_fanout < 0 || deg < (uint32_t)tail.next_fanout) ? deg : (uint32_t)tail.next_fanout;
        if (omit) {
          const uint32_t f = omit[__ldcg(source + k)];
          if (omit_value == 0xffffffffu ? (f != 0xffffffffu) : (f == omit_value)) c = 0;
        }
        s_dyn[k] = c;
      }
      __syncthreads();
      total = block_scan_array(s_dyn, S, s_warp);
      for (unsigned k = threadIdx.x; k <= S; k += blockDim.x) tail.next_col_off[k] = s_dyn[k];
      __syncthreads();
    }
    if (threadIdx.x == 0) {   // exactly CountOp::total
      next_meta->n_dst = ok ? S : 0u;
      next_meta->n_edges = total;
      next_meta->n_src = 0;
      next_meta->long_rows = 0;
      next_meta->err = !ok ? (err_in ? err_in : 2u) : (total > tail.next_cap_edges ? 1u : 0u);
      if (!ok) tail.next_col_off[0] = 0u;
    }
  }
  if (tail.row_offset && ok) {   // this layer's CSR row offsets from the finished use counts (RowOp)
    for (unsigned k = threadIdx.x; k < S; k += blockDim.x) s_dyn[k] = __ldcg(row_count + k);
    __syncthreads();
    block_scan_array(s_dyn, S, s_warp);
    for (unsigned k = threadIdx.x; k <= S; k += blockDim.x) tail.row_offset[k] = s_dyn[k];
  }
}

// CSR build step 1 with the row-offset scan done per block in shared memory (same results as k_scan<RowOp> + k_csr_fill)
__global__ void __launch_bounds__(FS_THREADS)
k_csr_fill_fused(const uint32_t *__restrict__ row_indices, const uint32_t *__restrict__ row_count, uint32_t *__restrict__ row_offset,
                 uint32_t *__restrict__ row_cursor, uint32_t *__restrict__ csr_tmp, const LayerMeta *meta) {
  extern __shared__ uint32_t s_dyn[];   // [n_src + 1]
  __shared__ unsigned s_warp[33];
  const unsigned S = meta->err ? 0u : meta->n_src;
  for (unsigned i = threadIdx.x; i < S; i += blockDim.x) s_dyn[i] = row_count[i];
  __syncthreads();
  block_scan_array(s_dyn, S, s_warp);
  if (blockIdx.x == 0)
    for (unsigned i = threadIdx.x; i <= S; i += blockDim.x) row_offset[i] = s_dyn[i];
  if (meta->err) return;
  const unsigned E = meta->n_edges;
  for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    const uint32_t s = row_indices[e];
    csr_tmp[s_dyn[s] + atomicAdd(&row_cursor[s], 1u)] = e;
  }
}

// Bottom layer: the packed index the gather-fused aggregation consumes (aggregate.cu, GATHERED): global source id of every edge
// with bit 31 set when this batch reads that source at least keep_min times (row_count = the finished per-source histogram).
__global__ void __launch_bounds__(256)
k_pack_gather_index(const uint32_t *__restrict__ sample_ans, const uint32_t *__restrict__ row_indices, const uint32_t *__restrict__ row_count,
                    uint32_t *__restrict__ gather_idx, const LayerMeta *meta, uint32_t keep_min) {
  if (meta->err) return;
  const unsigned E = meta->n_edges;
  for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x)
    gather_idx[e] = sample_ans[e] | (row_count[row_indices[e]] >= keep_min ? 0x80000000u : 0u);
}

// degrees from the CSC when the caller has none (clamped >= 1, core/graph.hpp:4525-4530)
__global__ void k_degrees_from_csc(const uint32_t *col_off, const uint32_t *row_idx, uint32_t *in_deg, uint32_t *out_deg,
                                   uint32_t V, uint64_t E, int phase) {
  uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
  if (phase == 0) {
    for (uint64_t v = tid; v < V; v += stride) { in_deg[v] = col_off[v + 1] - col_off[v]; out_deg[v] = 0; }
  } else if (phase == 1) {
    for (uint64_t e = tid; e < E; e += stride) atomicAdd(&out_deg[row_idx[e]], 1u);
  } else {
    for (uint64_t v = tid; v < V; v += stride) { if (in_deg[v] < 1) in_deg[v] = 1; if (out_deg[v] < 1) out_deg[v] = 1; }
  }
}

// =============================================================================================

extern "C" {

int nb_graph_create(nb_ctx *ctx, uint32_t n_vertices, uint64_t n_edges, const uint32_t *column_offset_host,
                    const uint32_t *row_indices_host, const uint32_t *in_degree_host, const uint32_t *out_degree_host,
                    nb_graph **out) {
  NB_REQUIRE(ctx && out && column_offset_host && (row_indices_host || n_edges == 0), NB_ERR_ARG, "nb_graph_create: NULL argument");
  NB_REQUIRE(n_vertices > 0 && n_edges < 0xffffffffull, NB_ERR_ARG, "nb_graph_create: |V| must be > 0 and |E| < 2^32 (u32 offsets)");
  NB_REQUIRE(column_offset_host[n_vertices] == (uint32_t)n_edges, NB_ERR_ARG, "column_offset[|V|] != |E|");
  NB_GUARD(ctx);
  nb_graph *g = new nb_graph();
  g->ctx = ctx; g->V = n_vertices; g->E = n_edges;
  g->col_off = g->row_idx = g->in_deg = g->out_deg = nullptr;
  NB_CUDA(cudaMalloc(&g->col_off, ((size_t)n_vertices + 1) * 4));
  NB_CUDA(cudaMalloc(&g->row_idx, (size_t)(n_edges ? n_edges : 1) * 4));
  NB_CUDA(cudaMalloc(&g->in_deg, (size_t)n_vertices * 4));
  NB_CUDA(cudaMalloc(&g->out_deg, (size_t)n_vertices * 4));
  NB_CUDA(cudaMemcpyAsync(g->col_off, column_offset_host, ((size_t)n_vertices + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (n_edges) NB_CUDA(cudaMemcpyAsync(g->row_idx, row_indices_host, (size_t)n_edges * 4, cudaMemcpyHostToDevice, ctx->stream));
  uint32_t mx = 0;
  for (uint32_t v = 0; v < n_vertices; v++) {
    uint32_t d = column_offset_host[v + 1] - column_offset_host[v];
    if (d > mx) mx = d;
  }
  g->max_in_degree = mx;
  if (in_degree_host && out_degree_host) {
    NB_CUDA(cudaMemcpyAsync(g->in_deg, in_degree_host, (size_t)n_vertices * 4, cudaMemcpyHostToDevice, ctx->stream));
    NB_CUDA(cudaMemcpyAsync(g->out_deg, out_degree_host, (size_t)n_vertices * 4, cudaMemcpyHostToDevice, ctx->stream));
  } else {
    for (int phase = 0; phase < 3; phase++) {
      k_degrees_from_csc<<<nb_grid(phase == 1 ? n_edges : n_vertices, 256), 256, 0, ctx->stream>>>(
          g->col_off, g->row_idx, g->in_deg, g->out_deg, n_vertices, n_edges, phase);
      NB_LAUNCH_CHECK(ctx);
    }
  }
  NB_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = g;
  return NB_OK;
}

__global__ void k_max_degree(const uint32_t *__restrict__ col_off, uint32_t V, uint32_t *out) {
  uint32_t m = 0;
  for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (uint64_t)gridDim.x * blockDim.x) m = max(m, col_off[v + 1] - col_off[v]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(FULL_MASK, m, o));
  if (lane_id() == 0 && m) atomicMax(out, m);
}

// The same from arrays that already live in device memory (a graph generated or loaded on the GPU): no host round trip.
int nb_graph_create_from_device(nb_ctx *ctx, uint32_t n_vertices, uint64_t n_edges, const uint32_t *column_offset_dev,
                                const uint32_t *row_indices_dev, nb_graph **out) {
  NB_REQUIRE(ctx && out && column_offset_dev && (row_indices_dev || n_edges == 0), NB_ERR_ARG, "nb_graph_create_from_device: NULL argument");
  NB_REQUIRE(n_vertices > 0 && n_edges < 0xffffffffull, NB_ERR_ARG, "nb_graph_create_from_device: |V| must be > 0 and |E| < 2^32 (u32 offsets)");
  NB_GUARD(ctx);
  nb_graph *g = new nb_graph();
  g->ctx = ctx; g->V = n_vertices; g->E = n_edges;
  g->col_off = g->row_idx = g->in_deg = g->out_deg = nullptr;
  auto fail = [&](int rc) { cudaFree(g->col_off); cudaFree(g->row_idx); cudaFree(g->in_deg); cudaFree(g->out_deg); delete g; return rc; };
#define NB_TRYG(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { nb_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); return fail(NB_ERR_CUDA); } } while (0)
  NB_TRYG(cudaMalloc(&g->col_off, ((size_t)n_vertices + 1) * 4));
  NB_TRYG(cudaMalloc(&g->row_idx, (size_t)(n_edges ? n_edges : 1) * 4));
  NB_TRYG(cudaMalloc(&g->in_deg, (size_t)n_vertices * 4));
  NB_TRYG(cudaMalloc(&g->out_deg, (size_t)n_vertices * 4));
  NB_TRYG(cudaMemcpyAsync(g->col_off, column_offset_dev, ((size_t)n_vertices + 1) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
  if (n_edges) NB_TRYG(cudaMemcpyAsync(g->row_idx, row_indices_dev, (size_t)n_edges * 4, cudaMemcpyDeviceToDevice, ctx->stream));
  uint32_t tail = 0, *mx_dev = nullptr;
  NB_TRYG(cudaMemcpyAsync(&tail, g->col_off + n_vertices, 4, cudaMemcpyDeviceToHost, ctx->stream));
  for (int phase = 0; phase < 3; phase++) {
    k_degrees_from_csc<<<nb_grid(phase == 1 ? n_edges : n_vertices, 256), 256, 0, ctx->stream>>>(g->col_off, g->row_idx, g->in_deg, g->out_deg, n_vertices, n_edges, phase);
    ctx->launches++;
  }
  NB_TRYG(cudaMalloc(&mx_dev, 4));
  cudaMemsetAsync(mx_dev, 0, 4, ctx->stream);
  k_max_degree<<<nb_grid(n_vertices, 256), 256, 0, ctx->stream>>>(g->col_off, n_vertices, mx_dev);
  ctx->launches++;
  uint32_t mx = 0;
  cudaMemcpyAsync(&mx, mx_dev, 4, cudaMemcpyDeviceToHost, ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  cudaFree(mx_dev);
  if (e != cudaSuccess) { nb_set_error("nb_graph_create_from_device: %s", cudaGetErrorString(e)); return fail(NB_ERR_CUDA); }
#undef NB_TRYG
  if (tail != (uint32_t)n_edges) { nb_set_error("column_offset[|V|] != |E|"); return fail(NB_ERR_ARG); }
  g->max_in_degree = mx;
  *out = g;
  return NB_OK;
}

int nb_graph_destroy(nb_graph *g) {
  if (!g) return NB_OK;
  DeviceGuard guard(g->ctx->device);
  cudaFree(g->col_off); cudaFree(g->row_idx); cudaFree(g->in_deg); cudaFree(g->out_deg);
  delete g;
  return NB_OK;
}

int nb_graph_info(nb_graph *g, uint32_t *n_vertices, uint64_t *n_edges, const uint32_t **column_offset_dev,
                  const uint32_t **row_indices_dev, const uint32_t **in_degree_dev, const uint32_t **out_degree_dev) {
  NB_REQUIRE(g, NB_ERR_ARG, "graph is NULL");
  if (n_vertices) *n_vertices = g->V;
  if (n_edges) *n_edges = g->E;
  if (column_offset_dev) *column_offset_dev = g->col_off;
  if (row_indices_dev) *row_indices_dev = g->row_idx;
  if (in_degree_dev) *in_degree_dev = g->in_deg;
  if (out_degree_dev) *out_degree_dev = g->out_deg;
  return NB_OK;
}

int nb_sampler_create(nb_ctx *ctx, nb_graph *g, int n_layers, const int *fanout, uint32_t max_batch, uint32_t flags,
                      uint64_t max_edges_hint, nb_sampler **out) {
  NB_REQUIRE(ctx && g && fanout && out, NB_ERR_ARG, "nb_sampler_create: NULL argument");
  NB_REQUIRE(n_layers >= 1 && n_layers <= NB_MAX_LAYERS, NB_ERR_ARG, "layers must be in [1,%d]", NB_MAX_LAYERS);
  NB_REQUIRE(max_batch >= 1, NB_ERR_ARG, "max_batch must be >= 1");
  for (int i = 0; i < n_layers; i++)
    NB_REQUIRE(fanout[i] == -1 || (fanout[i] >= 1 && fanout[i] <= 512), NB_ERR_UNSUPPORTED,
               "fanout[%d]=%d: supported values are -1 (all) and 1..512", i, fanout[i]);
  NB_GUARD(ctx);
  nb_sampler *s = new nb_sampler();
  memset(s, 0, sizeof(*s));
  s->ctx = ctx; s->g = g; s->L = n_layers; s->flags = flags; s->max_batch = max_batch;
  const bool merge = flags & NB_SAMPLER_MERGE_SRC_DST;
  // capacities: V_0 = batch, E_i <= V_i * f_i, S_i <= E_i (+V_i when merged), bounded by the graph
  uint64_t cap_dst = max_batch;  // seeds may repeat, so |V| does not bound layer 0
  size_t words = 0;              // arena size in 4-byte words
  auto take = [&](size_t n) { size_t at = words; words += (n + 31) & ~(size_t)31; return at; };
  struct Off { size_t destination, column_offset, sample_ans, row_indices, edge_dst, source, row_offset, row_count, row_cursor,
               column_indices, csr_tmp, csr_to_csc, long_rows, dst_local_id, src_to_dst, ewf, ewb, dst_base, dst_deg, gather_idx; } off[NB_MAX_LAYERS];
  uint64_t max_items = 0;
  for (int i = 0; i < n_layers; i++) {
    s->fanout[i] = fanout[i];
    uint64_t per = fanout[i] < 0 ? g->max_in_degree : (uint64_t)fanout[i];
    uint64_t cap_e = cap_dst * per;
    if (fanout[i] < 0 && max_edges_hint && max_edges_hint < cap_e) cap_e = max_edges_hint;
    if (i > 0 && cap_e > g->E) cap_e = g->E;  // dst of layers > 0 are unique vertices
    uint64_t cap_s = cap_e + (merge ? cap_dst : 0);
    if (cap_s > g->V) cap_s = g->V;
    if (cap_e >= 0x7fffffffull) {
      delete s;
      nb_set_error("layer %d edge capacity %llu exceeds 2^31", i, (unsigned long long)cap_e);
      return NB_ERR_CAPACITY;
    }
    LayerBuf &b = s->lay[i];
    b.cap_dst = (uint32_t)cap_dst; b.cap_edges = (uint32_t)cap_e; b.cap_src = (uint32_t)cap_s;
    Off &o = off[i];
    o.destination = i == 0 ? take(cap_dst) : 0;
    o.column_offset = take(cap_dst + 1);
    o.sample_ans = take(cap_e); o.row_indices = take(cap_e); o.edge_dst = take(cap_e);
    o.source = take(cap_s);
    o.row_offset = take(cap_s + 1); o.row_count = take(cap_s); o.row_cursor = take(cap_s);
    o.column_indices = take(cap_e); o.csr_tmp = take(cap_e); o.csr_to_csc = take(cap_e); o.long_rows = take(cap_s);
    o.dst_local_id = take(cap_dst); o.src_to_dst = take(cap_s);
    o.ewf = take(cap_e); o.ewb = take(cap_e);
    o.dst_base = take(cap_dst); o.dst_deg = take(cap_dst);
    o.gather_idx = (i == n_layers - 1 && g->V < 0x80000000u) ? take(cap_e) : (size_t)-1;
    if (cap_dst > max_items) max_items = cap_dst;
    if (cap_s > max_items) max_items = cap_s;
    cap_dst = cap_s;
  }
  s->n_words = (g->V + 31) / 32;
  if (s->n_words > max_items) max_items = s->n_words;
  s->max_tiles = (uint32_t)((max_items + SCAN_TILE - 1) / SCAN_TILE) + 1;
  size_t o_bitmap = take(s->n_words + 1), o_bitmap1 = take(s->n_words + 1), o_rank = take(s->n_words + 1);
  // two levels pay off when the marks are sparse: a batch touches at most cap_src level-0 words, so a flat pass over all of them only
  // wastes time when there are several times more words than that (papers100M: 3.47M words for <= 256K sources; products: 76K words,
  // nearly all touched -- flat is faster there, measured 0.17 vs 0.30 ms per step)
  uint64_t max_cap_src = 0;
  for (int i = 0; i < n_layers; i++) max_cap_src = max_cap_src > s->lay[i].cap_src ? max_cap_src : s->lay[i].cap_src;
  const bool two_level = g_sampler_two_level >= 0 ? (g_sampler_two_level == 1 && s->n_words > 64)
                                                  : ((uint64_t)s->n_words > 4 * max_cap_src && s->n_words > 32768);
  s->n_words_l1 = two_level ? (s->n_words + 31) / 32 : 0;
  size_t o_l1a = two_level ? take(s->n_words_l1 + 1) : 0, o_l1b = two_level ? take(s->n_words_l1 + 1) : 0;
  size_t o_state = take((size_t)s->max_tiles * 2 * 3 * n_layers);
  size_t o_meta = take((sizeof(LayerMeta) / 4) * (NB_MAX_LAYERS + 1));
  size_t o_params = take(sizeof(BatchParams) / 4 + 8);
  size_t bytes = words * 4;
  size_t free_b = 0, total_b = 0;
  NB_CUDA(cudaMemGetInfo(&free_b, &total_b));
  if (bytes > free_b) {
    delete s;
    nb_set_error("sampler arena needs %zu MiB, %zu MiB free", bytes >> 20, free_b >> 20);
    return NB_ERR_CAPACITY;
  }
  NB_CUDA(cudaMalloc(&s->arena, bytes));
  NB_CUDA(cudaMemsetAsync(s->arena, 0, bytes, ctx->stream));
  uint32_t *base = (uint32_t *)s->arena;
  for (int i = 0; i < n_layers; i++) {
    LayerBuf &b = s->lay[i];
    Off &o = off[i];
    b.destination = i == 0 ? base + o.destination : nullptr;
    b.column_offset = base + o.column_offset; b.sample_ans = base + o.sample_ans; b.row_indices = base + o.row_indices;
    b.edge_dst = base + o.edge_dst; b.source = base + o.source; b.row_offset = base + o.row_offset;
    b.row_count = base + o.row_count; b.row_cursor = base + o.row_cursor; b.column_indices = base + o.column_indices;
    b.csr_tmp = base + o.csr_tmp; b.csr_to_csc = base + o.csr_to_csc; b.long_rows = base + o.long_rows;
    b.dst_local_id = base + o.dst_local_id; b.src_to_dst = base + o.src_to_dst;
    b.ewf = (float *)(base + o.ewf); b.ewb = (float *)(base + o.ewb);
    b.dst_base = base + o.dst_base; b.dst_deg = base + o.dst_deg;
    b.gather_idx = o.gather_idx == (size_t)-1 ? nullptr : base + o.gather_idx;
  }
  for (int i = 1; i < n_layers; i++) s->lay[i].destination = s->lay[i - 1].source;  // layer chaining (FullyRepGraph.hpp:309)
  s->bitmap[0] = base + o_bitmap; s->bitmap[1] = base + o_bitmap1; s->word_rank = base + o_rank;
  s->bitmap_l1[0] = two_level ? base + o_l1a : nullptr; s->bitmap_l1[1] = two_level ? base + o_l1b : nullptr;
  s->tile_states = (unsigned long long *)(base + o_state);
  s->meta_dev = (LayerMeta *)(base + o_meta);
  s->params_dev = (BatchParams *)(base + o_params);
  NB_CUDA(cudaHostAlloc(&s->meta_ring, sizeof(LayerMeta) * (NB_MAX_LAYERS + 1) * nb_sampler::RING, cudaHostAllocDefault));
  memset(s->meta_ring, 0, sizeof(LayerMeta) * (NB_MAX_LAYERS + 1) * nb_sampler::RING);
  s->meta_host = s->meta_ring;
  s->meta_slot = 0;
  for (int r = 0; r < nb_sampler::RING; r++) {
    NB_CUDA(cudaHostAlloc(&s->stage[r], sizeof(BatchParams) + (size_t)max_batch * 4, cudaHostAllocDefault));
    NB_CUDA(cudaEventCreateWithFlags(&s->stage_done[r], cudaEventDisableTiming));
    NB_CUDA(cudaEventCreateWithFlags(&s->meta_ready[r], cudaEventDisableTiming));
  }
  for (int i = 0; i < NB_MAX_LAYERS; i++) {
    NB_CUDA(cudaEventCreateWithFlags(&s->ev_fork[i], cudaEventDisableTiming));
    NB_CUDA(cudaEventCreateWithFlags(&s->ev_join[i], cudaEventDisableTiming));
  }
  const char *ng = getenv("NB_NO_GRAPH");
  s->use_graph = !(ng && ng[0] == '1');
  if (g_sampler_fused < 0) { const char *e = getenv("NB_SAMPLER_FUSED"); g_sampler_fused = e ? atoi(e) : 0; }
  s->fused = g_sampler_fused;
  NB_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = s;
  return NB_OK;
}

int nb_sampler_destroy(nb_sampler *s) {
  if (!s) return NB_OK;
  DeviceGuard guard(s->ctx->device);
  cudaStreamSynchronize(s->ctx->stream);
  if (s->graph_exec) cudaGraphExecDestroy(s->graph_exec);
  for (int r = 0; r < nb_sampler::RING; r++) { cudaFreeHost(s->stage[r]); cudaEventDestroy(s->stage_done[r]); cudaEventDestroy(s->meta_ready[r]); }
  for (int i = 0; i < NB_MAX_LAYERS; i++) { cudaEventDestroy(s->ev_fork[i]); cudaEventDestroy(s->ev_join[i]); }
  cudaFree(s->arena);
  cudaFreeHost(s->meta_ring);
  delete s;
  return NB_OK;
}

static void fill_view(nb_sampler *s, int i, nb_layer_view *v) {
  const LayerBuf &b = s->lay[i];
  const LayerMeta &m = s->meta_host[i];
  const bool csr = (s->flags & NB_SAMPLER_BUILD_CSR) && !(i == s->L - 1 && s->L > 1 && (s->flags & NB_SAMPLER_NO_BOTTOM_CSR));
  const bool merge = s->flags & NB_SAMPLER_MERGE_SRC_DST;
  v->n_dst = m.n_dst; v->n_edges = m.n_edges; v->n_src = m.n_src; v->reserved = 0;
  v->destination = b.destination; v->column_offset = b.column_offset; v->sample_ans = b.sample_ans;
  v->row_indices = b.row_indices; v->source = b.source;
  v->row_offset = csr ? b.row_offset : nullptr; v->column_indices = csr ? b.column_indices : nullptr;
  v->csr_to_csc = csr ? b.csr_to_csc : nullptr;
  v->edge_weight_forward = b.ewf; v->edge_weight_backward = csr ? b.ewb : nullptr;
  v->dst_local_id = merge ? b.dst_local_id : nullptr; v->src_to_dst = merge ? b.src_to_dst : nullptr;
  v->source_use_count = b.row_count;
  v->gather_index = b.gather_idx;
}

// Every kernel of one mini-batch. All arguments are constants of the sampler (per-batch values come from
// params_dev), so the sequence is captured once into a CUDA graph and replayed.
// side != NULL (graph capture): a layer's CSR kernels depend only on that layer's relabel, so they are forked onto `side` and run
// beside the next layers' sampling; joined at the end. side == NULL: everything in order on st.
static int enqueue_kernels(nb_sampler *s, cudaStream_t st, cudaStream_t side = nullptr) {
  nb_ctx *ctx = s->ctx;
  nb_graph *g = s->g;
  const bool merge = s->flags & NB_SAMPLER_MERGE_SRC_DST, up = s->flags & NB_SAMPLER_UP_DEGREE,
             csr = s->flags & NB_SAMPLER_BUILD_CSR;
  const BatchParams *pp = s->params_dev;
  // with an odd number of layers the last layer leaves bitmap[0] marked, and layer 0 of the next batch uses it
  if (s->L & 1) {
    if (s->bitmap_l1[0]) {
      k_clear_two_level<<<sgrid(s->n_words_l1, 8, 8), 256, 0, st>>>(s->bitmap[0], s->bitmap_l1[0], s->n_words_l1);
      NB_LAUNCH_CHECK(ctx);
    } else NB_CUDA(cudaMemsetAsync(s->bitmap[0], 0, (size_t)(s->n_words + 1) * 4, st));
  }
  const unsigned fs_threads = (unsigned)g_sampler_block;
  bool have_col_off = false;   // this layer's column offsets + meta were left by the previous layer's relabel (RelabelTail)
  bool forked = false;
  for (int i = 0; i < s->L; i++) {
    LayerBuf &b = s->lay[i];
    LayerMeta *m = s->meta_dev + i;
    bool have_row_off = false;
    uint32_t *bm = s->bitmap[i & 1], *bm_other = (s->L > 1) ? s->bitmap[(i + 1) & 1] : nullptr;
    uint32_t *bm_l1 = s->bitmap_l1[i & 1], *bm_other_l1 = (s->L > 1) ? s->bitmap_l1[(i + 1) & 1] : nullptr;
    ScanWs ws0 = nb_scan_ws(s->tile_states + (size_t)(3 * i + 0) * s->max_tiles, s->max_tiles, pp),
           ws1 = nb_scan_ws(s->tile_states + (size_t)(3 * i + 1) * s->max_tiles, s->max_tiles, pp),
           ws2 = nb_scan_ws(s->tile_states + (size_t)(3 * i + 2) * s->max_tiles, s->max_tiles, pp);
    const bool layer_csr = csr && !(i == s->L - 1 && s->L > 1 && (s->flags & NB_SAMPLER_NO_BOTTOM_CSR));
    const int histogram = 1;   // per-source use counts: the CSR row lengths, UP_DEGREE's out-degrees and the aggregation's L2 hints
    const int bottom = i == s->L - 1 ? 1 : 0;
    uint32_t *next_base = i + 1 < s->L ? s->lay[i + 1].dst_base : nullptr, *next_deg = i + 1 < s->L ? s->lay[i + 1].dst_deg : nullptr;
    uint32_t *rc_ptr = (histogram || merge) ? b.row_count : nullptr;
    // ---- count + scan + neighbour selection
    uint32_t hs = 1; while (s->fanout[i] > 32 && hs < 2u * (uint32_t)s->fanout[i]) hs <<= 1;
    const int hash_slots = s->fanout[i] > 32 ? (int)hs : 0;
    const size_t smem_sample = ((size_t)((b.cap_dst + 1 + 31) & ~31u) + (size_t)hash_slots * (fs_threads / 32)) * 4;
    // general path with tails ("sampler_tail"): the sampling kernel's last block does the bitmap's popcount scan, the relabel
    // kernel's last block the next layer's count scan -- a layer is 2 kernels instead of 4, with the same look-back-free results
    const bool tails = g_sampler_tail && !s->fused;
    const bool rank_tail = tails && (g_sampler_tail & 1) && !bm_l1 && s->n_words <= FS_RANK_TAIL_MAX;
    uint32_t *tail_rank = rank_tail ? s->word_rank : nullptr;
    const unsigned sample_bps = g_sampler_bps > 0 ? (unsigned)g_sampler_bps : 8u;
    if (have_col_off) {
      launch_sample(st, b.cap_dst, s->fanout[i], g->col_off, g->row_idx, b.destination, b.column_offset, b.sample_ans, b.edge_dst, bm, m,
                    pp, (uint32_t)i, merge ? 1 : 0, rc_ptr, b.row_cursor, merge ? b.src_to_dst : nullptr, b.cap_src, bm_l1, b.dst_base, b.dst_deg,
                    sample_bps, tail_rank, s->n_words);
      NB_LAUNCH_CHECK(ctx);
    } else if ((s->fused && smem_sample <= FS_SMEM_MAX) || (tails && b.cap_dst <= 4096 && smem_sample <= 40 * 1024)) {
      // small-shape sampling kernel: every block scans the layer's counts in shared memory (layer 0 of the general path: 1024 seeds)
      const int group = (s->fanout[i] < 0 || s->fanout[i] > 16) ? 32 : (s->fanout[i] > 8 ? 16 : 8);
      const unsigned per_block = (fs_threads / 32) * (32 / group);
      unsigned grid = (b.cap_dst + per_block - 1) / per_block;
      const unsigned max_grid = (unsigned)ctx->sm_count * (g_sampler_bps > 0 ? (s->fused ? 1u : (unsigned)g_sampler_bps) : smem_sample <= 96 * 1024 ? 2u : 1u);
      if (grid > max_grid) grid = max_grid;
      if (grid < 1) grid = 1;
#define NB_FS(G)                                                                                                                   \
      do {                                                                                                                         \
        static bool attr = false;                                                                                                  \
        if (!attr) { NB_CUDA(cudaFuncSetAttribute(k_sample_fused<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FS_SMEM_MAX)); attr = true; } \
        k_sample_fused<G><<<grid, fs_threads, smem_sample, st>>>(g->col_off, g->row_idx, b.destination, b.dst_base, b.dst_deg, i > 0 ? 1 : 0,  \
            b.column_offset, b.sample_ans, b.edge_dst, bm, m, i ? m - 1 : nullptr, s->fanout[i], pp, (uint32_t)i, merge ? 1 : 0, bottom,   \
            hash_slots, rc_ptr, b.row_cursor, merge ? b.src_to_dst : nullptr, b.cap_src, b.cap_edges, b.cap_dst, bm_l1,                 \
            s->fused ? nullptr : tail_rank, s->n_words);                                                                            \
      } while (0)
      if (group == 32) NB_FS(32); else if (group == 16) NB_FS(16); else NB_FS(8);
#undef NB_FS
      NB_LAUNCH_CHECK(ctx);
    } else {
      // layers >= 1: the previous layer's relabel left every dst's adjacency base / degree next to the dst list (coalesced reads here
      // instead of two dependent random ones per dst)
      const uint32_t *dbase = i > 0 ? b.dst_base : nullptr, *ddeg = i > 0 ? b.dst_deg : nullptr;
      CountOp cop{g->col_off, b.destination, pp, b.column_offset, m, i ? m - 1 : nullptr, b.cap_edges, s->fanout[i], bottom, ddeg};
      k_scan<CountOp><<<sgrid(b.cap_dst, SCAN_TILE, 4), SCAN_THREADS, 0, st>>>(cop, ws0);
      NB_LAUNCH_CHECK(ctx);
      launch_sample(st, b.cap_dst, s->fanout[i], g->col_off, g->row_idx, b.destination, b.column_offset, b.sample_ans, b.edge_dst, bm, m,
                    pp, (uint32_t)i, merge ? 1 : 0, rc_ptr, b.row_cursor, merge ? b.src_to_dst : nullptr, b.cap_src, bm_l1, dbase, ddeg,
                    sample_bps, tail_rank, s->n_words);
      NB_LAUNCH_CHECK(ctx);
    }
    // ---- dedup ranks + source emission + relabel (+ histogram, weights)
    have_col_off = false;
    size_t smem_relabel = ((size_t)((s->n_words + 31) & ~31u) + s->n_words + 1) * 4;
    if (s->fused && !bm_l1 && smem_relabel <= FS_SMEM_MAX) {
      static bool attr = false;
      if (!attr) { NB_CUDA(cudaFuncSetAttribute(k_relabel_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FS_SMEM_MAX)); attr = true; }
      RelabelTail tail{};
      const size_t smem_tail = ((size_t)b.cap_src + 1) * 4;
      if ((g_sampler_tail & 2) && b.cap_src <= FS_TAIL_MAX && smem_tail <= FS_SMEM_MAX) {
        if (i + 1 < s->L) {
          tail.next_col_off = s->lay[i + 1].column_offset;
          tail.next_cap_edges = s->lay[i + 1].cap_edges;
          tail.next_fanout = s->fanout[i + 1];
          tail.next_bottom = i + 1 == s->L - 1 ? 1 : 0;
        }
        if (layer_csr) tail.row_offset = b.row_offset;
        tail.enabled = (tail.next_col_off || tail.row_offset) ? 1 : 0;
        if (tail.enabled && smem_tail > smem_relabel) smem_relabel = smem_tail;
      }
      const uint64_t work = (uint64_t)b.cap_edges / 4 + b.cap_dst + (uint64_t)s->n_words * 32;
      unsigned grid = (unsigned)((work + fs_threads - 1) / fs_threads);
      const unsigned per_sm = (unsigned)(FS_SMEM_MAX / smem_relabel);
      const unsigned bps_cap = g_sampler_bps > 0 ? (unsigned)g_sampler_bps : 4u;
      const unsigned max_grid = (unsigned)ctx->sm_count * (per_sm > bps_cap ? bps_cap : per_sm < 1u ? 1u : per_sm);
      if (grid > max_grid) grid = max_grid;
      k_relabel_fused<<<grid, fs_threads, smem_relabel, st>>>(
          b.sample_ans, b.row_indices, bm, b.row_count, b.destination, merge ? b.dst_local_id : nullptr, merge ? b.src_to_dst : nullptr,
          m, m + 1, histogram, up ? 0 : 1, b.ewf, b.edge_dst, b.column_offset, g->in_deg, g->out_deg, pp, b.source, s->n_words, bm_other,
          b.cap_src, g->col_off, next_base, next_deg, tail);
      NB_LAUNCH_CHECK(ctx);
      have_col_off = tail.next_col_off != nullptr;
      have_row_off = tail.row_offset != nullptr;
    } else {
      if (bm_l1) {
        Bitmap2Op bop{bm, bm_l1, s->word_rank, m, m + 1, s->n_words_l1, b.cap_src};
        k_scan<Bitmap2Op><<<sgrid(s->n_words_l1, SCAN_TILE, 4), SCAN_THREADS, 0, st>>>(bop, ws1);
        NB_LAUNCH_CHECK(ctx);
      } else if (!rank_tail) {
        BitmapOp bop{bm, s->word_rank, m, m + 1, s->n_words, b.cap_src};
        k_scan<BitmapOp><<<sgrid(s->n_words, SCAN_TILE, 4), SCAN_THREADS, 0, st>>>(bop, ws1);
        NB_LAUNCH_CHECK(ctx);
      }
      NextCount nc{};
      if (tails && (g_sampler_tail & 2) && i + 1 < s->L && b.cap_src <= FS_COUNT_TAIL_MAX) {
        nc.col_off = s->lay[i + 1].column_offset; nc.next_meta = m + 1; nc.cap_edges = s->lay[i + 1].cap_edges;
        nc.fanout = s->fanout[i + 1]; nc.bottom = i + 1 == s->L - 1 ? 1 : 0;
      }
      k_relabel<<<sgrid((uint64_t)b.cap_edges + b.cap_dst, 256, 8), 256, 0, st>>>(
          b.sample_ans, b.row_indices, bm, s->word_rank, b.row_count, b.destination, merge ? b.dst_local_id : nullptr,
          merge ? b.src_to_dst : nullptr, m, histogram, up ? 0 : 1, b.ewf, b.edge_dst, b.column_offset, g->in_deg, g->out_deg, pp,
          b.source, s->n_words, bm_other, g->col_off, next_base, next_deg, bm_l1, bm_other_l1, s->n_words_l1, nc);
      NB_LAUNCH_CHECK(ctx);
      have_col_off = nc.col_off != nullptr;
    }
    if (up) {
      k_weights_sampled<<<sgrid(b.cap_edges, 256, 8), 256, 0, st>>>(b.ewf, b.row_indices, b.edge_dst, b.column_offset, b.row_count, m, pp);
      NB_LAUNCH_CHECK(ctx);
    }
    if (b.gather_idx) {
      k_pack_gather_index<<<sgrid(b.cap_edges, 256, 8), 256, 0, st>>>(b.sample_ans, b.row_indices, b.row_count, b.gather_idx, m, (uint32_t)g_gather_keep_min);
      NB_LAUNCH_CHECK(ctx);
    }
    if (layer_csr) {
      cudaStream_t cst = st;
      if (side && i + 1 < s->L) {   // fork: the rest of this layer's work is off the chain the next layer waits for
        NB_CUDA(cudaEventRecord(s->ev_fork[i], st));
        NB_CUDA(cudaStreamWaitEvent(side, s->ev_fork[i], 0));
        cst = side;
        forked = true;
      }
      const size_t smem_csr = ((size_t)b.cap_src + 1) * 4;
      if (have_row_off) {
        k_csr_fill<<<sgrid(b.cap_edges, 256, 8), 256, 0, cst>>>(b.row_indices, b.row_offset, b.row_cursor, b.csr_tmp, m);
        NB_LAUNCH_CHECK(ctx);
        k_csr_rows<<<sgrid(b.cap_src, 256, 8), 256, 0, cst>>>(b.row_offset, b.csr_tmp, b.column_indices, b.csr_to_csc, b.edge_dst,
                                                                  b.ewb, b.ewf, b.long_rows, m, pp, 1);
        NB_LAUNCH_CHECK(ctx);
      } else if (s->fused && smem_csr <= FS_SMEM_MAX) {
        static bool attr = false;
        if (!attr) { NB_CUDA(cudaFuncSetAttribute(k_csr_fill_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FS_SMEM_MAX)); attr = true; }
        unsigned grid = (b.cap_edges + fs_threads - 1) / fs_threads;
        const unsigned max_grid = (unsigned)ctx->sm_count * (g_sampler_bps > 0 ? 1u : smem_csr <= 96 * 1024 ? 2u : 1u);
        if (grid > max_grid) grid = max_grid;
        if (grid < 1) grid = 1;
        k_csr_fill_fused<<<grid, fs_threads, smem_csr, cst>>>(b.row_indices, b.row_count, b.row_offset, b.row_cursor, b.csr_tmp, m);
        NB_LAUNCH_CHECK(ctx);
        k_csr_rows<<<sgrid(b.cap_src, 256, 8), 256, 0, cst>>>(b.row_offset, b.csr_tmp, b.column_indices, b.csr_to_csc, b.edge_dst,
                                                                 b.ewb, b.ewf, b.long_rows, m, pp, 1);
        NB_LAUNCH_CHECK(ctx);
      } else {
        RowOp rop{b.row_count, b.row_offset, m};
        k_scan<RowOp><<<sgrid(b.cap_src, SCAN_TILE, 4), SCAN_THREADS, 0, cst>>>(rop, ws2);
        NB_LAUNCH_CHECK(ctx);
        k_csr_fill<<<sgrid(b.cap_edges, 256, 8), 256, 0, cst>>>(b.row_indices, b.row_offset, b.row_cursor, b.csr_tmp, m);
        NB_LAUNCH_CHECK(ctx);
        k_csr_rows<<<sgrid(b.cap_src, 256, 8), 256, 0, cst>>>(b.row_offset, b.csr_tmp, b.column_indices, b.csr_to_csc, b.edge_dst,
                                                                 b.ewb, b.ewf, b.long_rows, m, pp, 1);
        NB_LAUNCH_CHECK(ctx);
      }
      if (cst != st) NB_CUDA(cudaEventRecord(s->ev_join[i], side));
    }
  }
  if (forked)
    for (int i = 0; i + 1 < s->L; i++)
      if (csr) NB_CUDA(cudaStreamWaitEvent(st, s->ev_join[i], 0));   // join: the batch is complete when st is
  return NB_OK;
}

// stage (params, seeds) in pinned memory, upload, run the batch (graph replay), fetch the sizes
static int run_batch(nb_sampler *s, const uint32_t *seeds, uint32_t n_seeds, int seeds_on_device, uint64_t rng_seed,
                     uint64_t rng_offset, int weight_type, const uint32_t *omit, uint32_t omit_value, int replay) {
  nb_ctx *ctx = s->ctx;
  cudaStream_t st = ctx->stream;
  const int slot = s->stage_next;
  s->stage_next = (slot + 1) % nb_sampler::RING;
  NB_CUDA(cudaEventSynchronize(s->stage_done[slot]));  // the upload that last used this slot has completed
  NB_CUDA(cudaEventSynchronize(s->meta_ready[slot]));  // ... and so has the copy of that batch's sizes into this slot of meta_ring
  BatchParams *hp = (BatchParams *)s->stage[slot];
  memset(hp, 0, sizeof(*hp));
  hp->rng_seed = rng_seed; hp->rng_offset = rng_offset; hp->omit = omit; hp->n_seeds = n_seeds;
  hp->weight_type = (uint32_t)weight_type; hp->omit_value = omit_value; hp->replay = (uint32_t)replay;
  hp->epoch = ++s->epoch;
  NB_CUDA(cudaMemcpyAsync(s->params_dev, hp, sizeof(BatchParams), cudaMemcpyHostToDevice, st));
  if (n_seeds) {
    if (seeds_on_device) {
      NB_CUDA(cudaMemcpyAsync(s->lay[0].destination, seeds, (size_t)n_seeds * 4, cudaMemcpyDeviceToDevice, st));
    } else {
      uint32_t *hs = (uint32_t *)(s->stage[slot] + sizeof(BatchParams));
      memcpy(hs, seeds, (size_t)n_seeds * 4);
      NB_CUDA(cudaMemcpyAsync(s->lay[0].destination, hs, (size_t)n_seeds * 4, cudaMemcpyHostToDevice, st));
    }
  }
  NB_CUDA(cudaEventRecord(s->stage_done[slot], st));
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  NB_CUDA(cudaStreamIsCapturing(st, &cap));
  if (s->use_graph && cap == cudaStreamCaptureStatusNone) {
    if (!s->graph_exec) {
      // the capture streams carry the priority of the stream the graph will be launched on: kernel nodes keep the priority of the
      // stream they were captured from
      cudaStream_t cst, side = nullptr;
      int prio = 0;
      if (g_sampler_capture_prio) NB_CUDA(cudaStreamGetPriority(st, &prio));
      NB_CUDA(cudaStreamCreateWithPriority(&cst, cudaStreamNonBlocking, prio));
      if (g_sampler_csr_branch) NB_CUDA(cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, prio));
      cudaGraph_t graph = nullptr;
      NB_CUDA(cudaStreamBeginCapture(cst, cudaStreamCaptureModeThreadLocal));
      const uint64_t launches0 = ctx->launches;
      int rc = enqueue_kernels(s, cst, side);
      cudaError_t ce = cudaStreamEndCapture(cst, &graph);
      cudaStreamDestroy(cst);
      if (side) cudaStreamDestroy(side);
      s->graph_kernels = ctx->launches - launches0;  // kernels per replay (what NB_LAUNCH_CHECK counted during capture)
      s->ctx->launches = launches0;
      if (rc != NB_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
      NB_CUDA(ce);
      NB_CUDA(cudaGraphInstantiate(&s->graph_exec, graph, 0));
      NB_CUDA(cudaGraphDestroy(graph));
    }
    NB_CUDA(cudaGraphLaunch(s->graph_exec, st));
    ctx->launches += s->graph_kernels;
  } else {
    int rc = enqueue_kernels(s, st);
    if (rc != NB_OK) return rc;
  }
  NB_CUDA(cudaMemcpyAsync(s->meta_ring + (size_t)slot * (NB_MAX_LAYERS + 1), s->meta_dev, sizeof(LayerMeta) * (s->L + 1), cudaMemcpyDeviceToHost, st));
  NB_CUDA(cudaEventRecord(s->meta_ready[slot], st));
  s->meta_slot = slot;
  return NB_OK;
}

static int finish_batch(nb_sampler *s, nb_layer_view *views_out) {
  // the batch enqueued last (an earlier one's arena contents are already being overwritten by it, so its sizes are the only
  // meaningful ones); waits for that batch only, not for later work on the stream
  NB_CUDA(cudaEventSynchronize(s->meta_ready[s->meta_slot]));
  s->meta_host = s->meta_ring + (size_t)s->meta_slot * (NB_MAX_LAYERS + 1);
  for (int i = 0; i < s->L; i++) {
    if (s->meta_host[i].err) {
      nb_set_error("sampler layer %d: %s capacity exceeded (E=%u cap %u, S=%u cap %u)", i,
                   s->meta_host[i].err == 1 ? "edge" : "source", s->meta_host[i].n_edges, s->lay[i].cap_edges,
                   s->meta_host[i].n_src, s->lay[i].cap_src);
      return NB_ERR_CAPACITY;
    }
    if (views_out) fill_view(s, i, views_out + i);
  }
  return NB_OK;
}

int nb_sampler_sample(nb_sampler *s, const uint32_t *seeds, uint32_t n_seeds, int seeds_on_device, uint64_t rng_seed,
                      uint64_t rng_offset, int weight_type, const uint32_t *omit_flag_dev, uint32_t omit_value,
                      nb_layer_view *views_out, int sync) {
  NB_REQUIRE(s && (seeds || n_seeds == 0), NB_ERR_ARG, "nb_sampler_sample: NULL argument");
  NB_REQUIRE(n_seeds <= s->lay[0].cap_dst, NB_ERR_CAPACITY, "batch of %u seeds exceeds the sampler's max_batch %u", n_seeds, s->lay[0].cap_dst);
  NB_REQUIRE(weight_type >= 0 && weight_type <= 3, NB_ERR_ARG, "bad weight_type %d", weight_type);
  NB_GUARD(s->ctx);
  int rc = run_batch(s, seeds, n_seeds, seeds_on_device, rng_seed, rng_offset, weight_type, omit_flag_dev, omit_value, 0);
  if (rc != NB_OK) return rc;
  if (sync) return finish_batch(s, views_out);
  return NB_OK;
}

// Replay: the recorded draws of every layer are uploaded first; the kernels then only mark and relabel them.
int nb_sampler_replay(nb_sampler *s, const uint32_t *seeds_host, uint32_t n_seeds, const uint32_t *const *sample_ans_host,
                      const uint32_t *n_edges_host, int weight_type, nb_layer_view *views_out) {
  NB_REQUIRE(s && seeds_host && sample_ans_host && n_edges_host, NB_ERR_ARG, "nb_sampler_replay: NULL argument");
  NB_REQUIRE(n_seeds <= s->lay[0].cap_dst, NB_ERR_CAPACITY, "batch of %u seeds exceeds max_batch", n_seeds);
  NB_GUARD(s->ctx);
  for (int i = 0; i < s->L; i++) {
    NB_REQUIRE(n_edges_host[i] <= s->lay[i].cap_edges, NB_ERR_CAPACITY, "replay layer %d: %u edges exceed capacity %u", i, n_edges_host[i], s->lay[i].cap_edges);
    if (n_edges_host[i])
      NB_CUDA(cudaMemcpyAsync(s->lay[i].sample_ans, sample_ans_host[i], (size_t)n_edges_host[i] * 4, cudaMemcpyHostToDevice, s->ctx->stream));
  }
  int rc = run_batch(s, seeds_host, n_seeds, 0, 0, 0, weight_type, nullptr, 0, 1);
  if (rc != NB_OK) return rc;
  rc = finish_batch(s, views_out);
  if (rc != NB_OK) return rc;
  for (int i = 0; i < s->L; i++)
    NB_REQUIRE(s->meta_host[i].n_edges == n_edges_host[i], NB_ERR_ARG, "replay layer %d: supplied %u edges, column offsets total %u", i,
               n_edges_host[i], s->meta_host[i].n_edges);
  return NB_OK;
}

// Completes an nb_sampler_sample(..., sync = 0): blocks until that batch's sizes are on the host, reports arena
// overflow, fills the views. Lets a caller overlap the sampling of batch i+1 with its work on batch i.
int nb_sampler_wait(nb_sampler *s, nb_layer_view *views_out) {
  NB_REQUIRE(s, NB_ERR_ARG, "nb_sampler_wait: NULL sampler");
  NB_GUARD(s->ctx);
  return finish_batch(s, views_out);
}

int nb_sampler_layer(nb_sampler *s, int layer, nb_layer_view *out) {
  NB_REQUIRE(s && out && layer >= 0 && layer < s->L, NB_ERR_ARG, "nb_sampler_layer: bad argument");
  NB_REQUIRE(s->meta_host[layer].err == 0, NB_ERR_CAPACITY, "layer %d overflowed its arena", layer);
  fill_view(s, layer, out);
  return NB_OK;
}

// Device addresses of a layer's sizes, for the *_dyn entry points that take their extents from device memory
// (no host round trip between sampling and the kernels that consume the sampled layer).
int nb_sampler_sizes_dev(nb_sampler *s, int layer, const uint32_t **n_dst_dev, const uint32_t **n_edges_dev,
                         const uint32_t **n_src_dev, uint32_t *cap_dst, uint32_t *cap_edges, uint32_t *cap_src) {
  NB_REQUIRE(s && layer >= 0 && layer < s->L, NB_ERR_ARG, "nb_sampler_sizes_dev: bad argument");
  LayerMeta *m = s->meta_dev + layer;
  if (n_dst_dev) *n_dst_dev = &m->n_dst;
  if (n_edges_dev) *n_edges_dev = &m->n_edges;
  if (n_src_dev) *n_src_dev = &m->n_src;
  if (cap_dst) *cap_dst = s->lay[layer].cap_dst;
  if (cap_edges) *cap_edges = s->lay[layer].cap_edges;
  if (cap_src) *cap_src = s->lay[layer].cap_src;
  return NB_OK;
}


// =============================================================================================
// Stage-by-stage entry points with the reference's own call shapes, for unmodified callers of
// Cuda_Stream::sample_processing_* (core/FullyRepGraph.hpp:326-524 drives them one stage at a time
// with host round trips in between). They run the same kernels as nb_sampler_sample on caller-owned
// arrays; transient state (bitmap, ranks, scan tiles) lives in the ctx scratch buffer.
struct LegacyState {
  LayerMeta meta[2];
  BatchParams params;
};

__global__ void k_update_ri(uint32_t *r_i, const uint32_t *__restrict__ src_index, uint32_t n) {
  for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) r_i[e] = src_index[r_i[e]];
}
__global__ void k_map_ids(uint32_t *out, const uint32_t *__restrict__ ids, const uint32_t *__restrict__ map, uint32_t n) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = map[ids[i]];
}
// in_degree[dst] = sampled column length (fanout when 0 and cache_fanout > 0), out_degree[src]++ per sampled edge
// (up_date_degree / update_cache_degree, cuda/ntsCUDATransferKernel.cuh:238-292)
__global__ void k_update_degree(uint32_t *out_degree, uint32_t *in_degree, uint32_t n_dst, const uint32_t *__restrict__ destination,
                                const uint32_t *__restrict__ source, const uint32_t *__restrict__ col_off,
                                const uint32_t *__restrict__ row_indices, int cache_fanout) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_dst; i += gridDim.x * blockDim.x) {
    const uint32_t b = col_off[i], e = col_off[i + 1];
    uint32_t len = e - b;
    if (len == 0 && cache_fanout > 0) len = (uint32_t)cache_fanout;
    in_degree[destination[i]] = len;
    for (uint32_t j = b; j < e; j++) atomicAdd(&out_degree[source[row_indices[j]]], 1u);
  }
}
// get_weight / get_mean_weight (ibid. :294-342) on caller-owned |V| degree arrays
__global__ void k_legacy_weight(float *edge_weight, const uint32_t *__restrict__ out_degree, const uint32_t *__restrict__ in_degree,
                                uint32_t n_dst, const uint32_t *__restrict__ destination, const uint32_t *__restrict__ source,
                                const uint32_t *__restrict__ col_off, const uint32_t *__restrict__ row_indices, int mean) {
  const unsigned lane = lane_id(), warps = (gridDim.x * blockDim.x) >> 5;
  for (unsigned d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; d < n_dst; d += warps) {
    const uint32_t b = col_off[d], e = col_off[d + 1], id = in_degree[destination[d]];
    for (uint32_t j = b + lane; j < e; j += 32)
      edge_weight[j] = edge_weight_fn(out_degree[source[row_indices[j]]], id, e - b, mean ? NB_WEIGHT_MEAN_SAMPLED : NB_WEIGHT_SUM);
  }
}

static int legacy_state(nb_ctx *ctx, uint32_t n_items_for_scan, uint32_t n_vertices, uint64_t n_edges, LegacyState **st,
                        unsigned long long **tiles, uint32_t **bitmap, uint32_t **word_rank, uint32_t **edge_dst, size_t *n_tiles_out) {
  const uint32_t n_words = (n_vertices + 31) / 32;
  uint32_t items = n_items_for_scan > n_words ? n_items_for_scan : n_words;
  const size_t n_tiles = (items + SCAN_TILE - 1) / SCAN_TILE + 1;
  size_t bytes = 1024 + n_tiles * 8 + ((size_t)n_words + 64) * 8 + (size_t)n_edges * 4 + 256;
  uint8_t *p;
  int rc = nb_ctx_scratch(ctx, bytes, (void **)&p);
  if (rc) return rc;
  *st = (LegacyState *)p;
  *tiles = (unsigned long long *)(p + 1024);
  *bitmap = (uint32_t *)(p + 1024 + n_tiles * 8);
  *word_rank = *bitmap + n_words + 32;
  *edge_dst = *word_rank + n_words + 32;
  NB_CUDA(cudaMemsetAsync(*tiles, 0, n_tiles * 8, ctx->stream));
  *n_tiles_out = n_tiles;
  return NB_OK;
}

int nb_sample_count(nb_ctx *ctx, const uint32_t *dst_dev, uint32_t *local_column_offset_dev, const uint32_t *global_column_offset_dev,
                    uint32_t dst_size, uint32_t fanout, const uint32_t *omit_flag_dev, uint32_t omit_value, uint32_t *edge_size_out) {
  NB_REQUIRE(ctx && local_column_offset_dev && global_column_offset_dev && edge_size_out && (dst_dev || dst_size == 0), NB_ERR_ARG,
             "nb_sample_count: NULL argument");
  NB_GUARD(ctx);
  LegacyState *st; unsigned long long *tiles; uint32_t *bitmap, *rank, *edge_dst;
  size_t n_tiles_ws = 0;
  int rc = legacy_state(ctx, dst_size, 32, 0, &st, &tiles, &bitmap, &rank, &edge_dst, &n_tiles_ws);
  if (rc) return rc;
  LegacyState h;
  memset(&h, 0, sizeof(h));
  h.params.n_seeds = dst_size;
  h.params.omit = omit_flag_dev; h.params.omit_value = omit_value; h.params.epoch = 1;
  NB_CUDA(cudaMemcpyAsync(st, &h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
  CountOp cop{global_column_offset_dev, dst_dev, &st->params, local_column_offset_dev, &st->meta[0], nullptr, 0xffffffffu, (int)fanout, 1};
  ScanWs ws = nb_scan_ws(tiles, n_tiles_ws, &st->params);
  k_scan<CountOp><<<nb_grid(dst_size, SCAN_TILE, 4), SCAN_THREADS, 0, ctx->stream>>>(cop, ws);
  NB_LAUNCH_CHECK(ctx);
  NB_CUDA(cudaMemcpyAsync(edge_size_out, &st->meta[0].n_edges, 4, cudaMemcpyDeviceToHost, ctx->stream));
  NB_CUDA(cudaStreamSynchronize(ctx->stream));  // the reference returns edge_size by reference
  return NB_OK;
}

int nb_sample_traverse(nb_ctx *ctx, const uint32_t *destination_dev, const uint32_t *column_offset_dev, uint32_t *r_i_dev,
                       const uint32_t *global_column_offset_dev, const uint32_t *global_row_indices_dev, uint32_t *src_index_dev,
                       uint32_t vtx_size, uint32_t edge_size, uint32_t n_vertices, uint32_t *src_dev, uint32_t *src_count_dev,
                       uint32_t layer, uint32_t fanout, int add_dst_to_src, uint64_t rng_seed, uint64_t rng_offset) {
  NB_REQUIRE(ctx && column_offset_dev && global_column_offset_dev && global_row_indices_dev && src_index_dev && src_dev && src_count_dev,
             NB_ERR_ARG, "nb_sample_traverse: NULL argument");
  NB_REQUIRE(fanout >= 1 && fanout <= 512, NB_ERR_UNSUPPORTED, "fanout %u: supported values are 1..512", fanout);
  NB_GUARD(ctx);
  global_row_indices_dev = (const uint32_t *)nb_mirror_host(ctx, global_row_indices_dev, 1);  // adjacency left in pinned host memory by the caller
  LegacyState *st; unsigned long long *tiles; uint32_t *bitmap, *rank, *edge_dst;
  size_t n_tiles_ws = 0;
  int rc = legacy_state(ctx, vtx_size, n_vertices, edge_size, &st, &tiles, &bitmap, &rank, &edge_dst, &n_tiles_ws);
  if (rc) return rc;
  const uint32_t n_words = (n_vertices + 31) / 32;
  LegacyState h;
  memset(&h, 0, sizeof(h));
  h.meta[0].n_dst = vtx_size; h.meta[0].n_edges = edge_size;
  h.params.rng_seed = rng_seed; h.params.rng_offset = rng_offset; h.params.epoch = 1;
  NB_CUDA(cudaMemcpyAsync(st, &h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
  NB_CUDA(cudaMemsetAsync(bitmap, 0, (size_t)(n_words + 1) * 4, ctx->stream));
  launch_sample(ctx->stream, vtx_size, (int)fanout, global_column_offset_dev, global_row_indices_dev, destination_dev, column_offset_dev,
                r_i_dev, edge_dst, bitmap, &st->meta[0], &st->params, layer, add_dst_to_src ? 1 : 0);
  NB_LAUNCH_CHECK(ctx);
  BitmapOp bop{bitmap, rank, &st->meta[0], &st->meta[1], n_words, 0xffffffffu};
  ScanWs ws = nb_scan_ws(tiles, n_tiles_ws, &st->params);
  k_scan<BitmapOp><<<nb_grid(n_words, SCAN_TILE, 4), SCAN_THREADS, 0, ctx->stream>>>(bop, ws);
  NB_LAUNCH_CHECK(ctx);
  k_emit_sources<<<nb_grid(n_words, 8, 8), 256, 0, ctx->stream>>>(bitmap, rank, src_dev, nullptr, nullptr, nullptr, &st->meta[0], n_words, src_index_dev);
  NB_LAUNCH_CHECK(ctx);
  NB_CUDA(cudaMemcpyAsync(src_count_dev, &st->meta[0].n_src, 4, cudaMemcpyDeviceToDevice, ctx->stream));
  return NB_OK;
}

int nb_sample_update_ri(nb_ctx *ctx, uint32_t *r_i_dev, const uint32_t *src_index_dev, uint32_t edge_size) {
  NB_REQUIRE(ctx && (edge_size == 0 || (r_i_dev && src_index_dev)), NB_ERR_ARG, "nb_sample_update_ri: NULL argument");
  NB_GUARD(ctx);
  if (!edge_size) return NB_OK;
  k_update_ri<<<nb_grid(edge_size, 256, 8), 256, 0, ctx->stream>>>(r_i_dev, src_index_dev, edge_size);
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
}

int nb_set_dst_local_index(nb_ctx *ctx, const uint32_t *src_index_dev, const uint32_t *destination_dev, uint32_t n_dst, uint32_t *dst_to_local_dev) {
  NB_REQUIRE(ctx && (n_dst == 0 || (src_index_dev && destination_dev && dst_to_local_dev)), NB_ERR_ARG, "nb_set_dst_local_index: NULL argument");
  NB_GUARD(ctx);
  if (!n_dst) return NB_OK;
  k_map_ids<<<nb_grid(n_dst, 256, 8), 256, 0, ctx->stream>>>(dst_to_local_dev, destination_dev, src_index_dev, n_dst);
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
}

int nb_update_degree(nb_ctx *ctx, uint32_t *out_degree_dev, uint32_t *in_degree_dev, uint32_t n_vertices, uint32_t n_dst,
                     const uint32_t *destination_dev, const uint32_t *source_dev, const uint32_t *column_offset_dev,
                     const uint32_t *row_indices_dev, int cache_fanout) {
  NB_REQUIRE(ctx && out_degree_dev && in_degree_dev, NB_ERR_ARG, "nb_update_degree: NULL argument");
  NB_GUARD(ctx);
  if (n_vertices) {  // ReFreshDegree
    NB_CUDA(cudaMemsetAsync(out_degree_dev, 0, (size_t)n_vertices * 4, ctx->stream));
    NB_CUDA(cudaMemsetAsync(in_degree_dev, 0, (size_t)n_vertices * 4, ctx->stream));
  }
  if (!n_dst) return NB_OK;
  k_update_degree<<<nb_grid(n_dst, 256, 8), 256, 0, ctx->stream>>>(out_degree_dev, in_degree_dev, n_dst, destination_dev, source_dev,
                                                                   column_offset_dev, row_indices_dev, cache_fanout);
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
}

int nb_edge_weight(nb_ctx *ctx, float *edge_weight_dev, const uint32_t *out_degree_dev, const uint32_t *in_degree_dev, uint32_t n_dst,
                   const uint32_t *destination_dev, const uint32_t *source_dev, const uint32_t *column_offset_dev,
                   const uint32_t *row_indices_dev, int mean) {
  NB_REQUIRE(ctx && (n_dst == 0 || (edge_weight_dev && out_degree_dev && in_degree_dev && destination_dev && source_dev && column_offset_dev)),
             NB_ERR_ARG, "nb_edge_weight: NULL argument");
  NB_GUARD(ctx);
  if (!n_dst) return NB_OK;
  k_legacy_weight<<<nb_grid(n_dst, 8, 8), 256, 0, ctx->stream>>>(edge_weight_dev, out_degree_dev, in_degree_dev, n_dst, destination_dev,
                                                                 source_dev, column_offset_dev, row_indices_dev, mean);
  NB_LAUNCH_CHECK(ctx);
  return NB_OK;
}

}  // extern "C"
