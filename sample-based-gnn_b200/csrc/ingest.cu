// ingest.cu -- global in-edge CSC built on the device from the raw edge list.
//
// Replaces FullyRepGraph::GenerateAll / ReadRepGraphFromRawFile (core/FullyRepGraph.hpp:724-798): a two-pass counting sort on
// the host -- column = dst, entries of a column in FILE ORDER -- followed by the degree arrays (core/graph.hpp:4525-4530,
// clamped >= 1). SURVEY section 8 row a1 / (f)3. The order inside a column is part of the parity contract (take-all columns
// return the neighbours "in stored order"), so the device build has to be a STABLE sort by dst:
//   k_split_pairs   (src,dst) pairs -> keys = dst, vals = src, in/out degree histograms (atomics; counts only, order-free)
//   per 8-bit digit of dst, least significant first (LSD radix sort, stable in every pass):
//     k_rs_hist     per tile of 2048 edges: digit histogram                         -> tile_hist[digit][tile]
//     k_scan        exclusive scan of tile_hist in digit-major order                -> first output slot of (digit, tile)
//     k_rs_scatter  stable scatter: a tile is 64 warp-sized segments in element order; __match_any gives each element its
//                   rank among the equal digits of its segment, a column scan over the 64 segment histograms gives the
//                   segment's offset inside the tile
//   k_scan          exclusive scan of the in-degree histogram -> column_offset
// HBM bound: 16 bytes read + 16 written per edge and pass (3 passes for |V| < 2^24): ~1 ms per 100M edges at HBM speed, against
// seconds for the host sort. No library sort is used.
#include "common.cuh"
#include "scan.cuh"

constexpr int RS_THREADS = 256, RS_ITEMS = 8, RS_TILE = RS_THREADS * RS_ITEMS, RS_BINS = 256, RS_SEGS = RS_TILE / 32;

struct FlatScanOp {
  const uint32_t *in;
  uint32_t *out, *total_out;
  uint32_t n_items;
  __device__ unsigned n() const { return n_items; }
  __device__ unsigned load(unsigned i) const { return in[i]; }
  __device__ void store(unsigned i, unsigned excl, unsigned) const { out[i] = excl; }
  __device__ void total(unsigned t) const { if (total_out) *total_out = t; }
};

__global__ void __launch_bounds__(256)
k_split_pairs(const uint32_t *__restrict__ pairs, uint64_t n_edges, uint32_t V, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals,
              uint32_t *__restrict__ in_cnt, uint32_t *__restrict__ out_cnt, uint32_t *__restrict__ bad) {
  const uint2 *p2 = reinterpret_cast<const uint2 *>(pairs);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_edges; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint2 e = p2[i];  // (src, dst)
    if (e.x >= V || e.y >= V) { atomicAdd(bad, 1u); keys[i] = 0; vals[i] = 0; continue; }
    keys[i] = e.y;
    vals[i] = e.x;
    atomicAdd(&in_cnt[e.y], 1u);
    atomicAdd(&out_cnt[e.x], 1u);
  }
}

__global__ void __launch_bounds__(RS_THREADS)
k_rs_hist(const uint32_t *__restrict__ keys, uint64_t n, int shift, uint32_t n_tiles, uint32_t *__restrict__ tile_hist) {
  __shared__ uint32_t h[RS_BINS];
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)tile * RS_TILE;
#pragma unroll
    for (int k = 0; k < RS_ITEMS; k++) {
      const uint64_t i = base + (uint64_t)k * RS_THREADS + threadIdx.x;
      if (i < n) atomicAdd(&h[(keys[i] >> shift) & (RS_BINS - 1)], 1u);
    }
    __syncthreads();
    tile_hist[(uint64_t)threadIdx.x * n_tiles + tile] = h[threadIdx.x];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(RS_THREADS)
k_rs_scatter(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint64_t n, int shift, uint32_t n_tiles,
             const uint32_t *__restrict__ tile_base, uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out) {
  __shared__ uint16_t seg[RS_SEGS][RS_BINS];  // count, then exclusive prefix over the segments, per digit
  __shared__ uint32_t gbase[RS_BINS];
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    uint32_t *z = reinterpret_cast<uint32_t *>(&seg[0][0]);
    for (unsigned i = threadIdx.x; i < RS_SEGS * RS_BINS / 2; i += RS_THREADS) z[i] = 0;
    gbase[threadIdx.x] = tile_base[(uint64_t)threadIdx.x * n_tiles + tile];
    __syncthreads();
    const uint64_t base = (uint64_t)tile * RS_TILE;
    uint32_t key[RS_ITEMS], val[RS_ITEMS], rank[RS_ITEMS];
#pragma unroll
    for (int k = 0; k < RS_ITEMS; k++) {
      const uint64_t i = base + (uint64_t)k * RS_THREADS + threadIdx.x;   // element order = (k, warp, lane)
      const bool valid = i < n;
      key[k] = valid ? keys[i] : 0u;
      val[k] = valid ? vals[i] : 0u;
      const uint32_t d = valid ? ((key[k] >> shift) & (RS_BINS - 1)) : RS_BINS;
      const unsigned m = __match_any_sync(FULL_MASK, d);
      rank[k] = __popc(m & lt);
      if (valid && rank[k] == 0) seg[k * (RS_THREADS / 32) + warp][d] = (uint16_t)__popc(m);
    }
    __syncthreads();
    {  // thread d: exclusive prefix of digit d over the 64 segments, in element order
      uint32_t run = 0;
#pragma unroll 8
      for (int s = 0; s < RS_SEGS; s++) {
        const uint32_t c = seg[s][threadIdx.x];
        seg[s][threadIdx.x] = (uint16_t)run;
        run += c;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RS_ITEMS; k++) {
      const uint64_t i = base + (uint64_t)k * RS_THREADS + threadIdx.x;
      if (i < n) {
        const uint32_t d = (key[k] >> shift) & (RS_BINS - 1);
        const uint32_t pos = gbase[d] + seg[k * (RS_THREADS / 32) + warp][d] + rank[k];
        keys_out[pos] = key[k];
        vals_out[pos] = val[k];
      }
    }
    __syncthreads();
  }
}

__global__ void k_finish_degrees(uint32_t *__restrict__ in_deg, uint32_t *__restrict__ out_deg, uint32_t V, uint32_t *__restrict__ max_in) {
  uint32_t mx = 0;
  for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
    const uint32_t d = in_deg[v];
    mx = max(mx, d);
    if (d < 1) in_deg[v] = 1;           // Graph::in/out_degree_for_backward are clamped >= 1 (core/graph.hpp:4525-4530)
    if (out_deg[v] < 1) out_deg[v] = 1;
  }
  mx = max(mx, __shfl_xor_sync(FULL_MASK, mx, 16)); mx = max(mx, __shfl_xor_sync(FULL_MASK, mx, 8));
  mx = max(mx, __shfl_xor_sync(FULL_MASK, mx, 4));  mx = max(mx, __shfl_xor_sync(FULL_MASK, mx, 2));
  mx = max(mx, __shfl_xor_sync(FULL_MASK, mx, 1));
  if ((threadIdx.x & 31) == 0 && mx) atomicMax(max_in, mx);
}

struct IngestTmp {
  uint32_t *pairs = nullptr, *keys[2] = {nullptr, nullptr}, *vals = nullptr, *tile_hist = nullptr, *small = nullptr;
  unsigned long long *tile_state = nullptr;
  BatchParams *params = nullptr;
  ~IngestTmp() {
    cudaFree(pairs); cudaFree(keys[0]); cudaFree(keys[1]); cudaFree(vals); cudaFree(tile_hist); cudaFree(small);
    cudaFree(tile_state); cudaFree(params);
  }
};

extern "C" int nb_graph_create_from_pairs(nb_ctx *ctx, uint32_t n_vertices, uint64_t n_edges, const uint32_t *pairs,
                                          int pairs_on_device, nb_graph **out) {
  NB_REQUIRE(ctx && out && (pairs || n_edges == 0), NB_ERR_ARG, "nb_graph_create_from_pairs: NULL argument");
  NB_REQUIRE(n_vertices > 0 && n_edges < 0xffffffffull, NB_ERR_ARG, "nb_graph_create_from_pairs: |V| must be > 0 and |E| < 2^32 (u32 offsets)");
  NB_GUARD(ctx);
  cudaStream_t st = ctx->stream;
  const uint32_t V = n_vertices;
  const uint64_t E = n_edges, En = E ? E : 1;
  const uint32_t n_tiles = (uint32_t)((En + RS_TILE - 1) / RS_TILE);
  const uint64_t n_hist = (uint64_t)RS_BINS * n_tiles;
  NB_REQUIRE(n_hist < 0xffffffffull, NB_ERR_UNSUPPORTED, "nb_graph_create_from_pairs: edge list too long for the tile histogram");
  NB_REQUIRE(pairs_on_device == 0 || ((uintptr_t)pairs & 7) == 0, NB_ERR_ARG, "nb_graph_create_from_pairs: device pairs must be 8-byte aligned");
  nb_graph *g = new nb_graph();
  g->ctx = ctx; g->V = V; g->E = E;
  g->col_off = g->row_idx = g->in_deg = g->out_deg = nullptr;
  g->max_in_degree = 0;
  IngestTmp t;
  auto fail = [&](int rc) { cudaFree(g->col_off); cudaFree(g->row_idx); cudaFree(g->in_deg); cudaFree(g->out_deg); delete g; return rc; };
#define NB_TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { nb_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); return fail(NB_ERR_CUDA); } } while (0)
  NB_TRY(cudaMalloc(&g->col_off, ((size_t)V + 1) * 4));
  NB_TRY(cudaMalloc(&g->row_idx, En * 4));
  NB_TRY(cudaMalloc(&g->in_deg, (size_t)V * 4));
  NB_TRY(cudaMalloc(&g->out_deg, (size_t)V * 4));
  NB_TRY(cudaMemsetAsync(g->in_deg, 0, (size_t)V * 4, st));
  NB_TRY(cudaMemsetAsync(g->out_deg, 0, (size_t)V * 4, st));
  const uint32_t scan_tiles = (uint32_t)((max(n_hist, (uint64_t)V) + SCAN_TILE - 1) / SCAN_TILE) + 1;
  int bits = 0;
  while (bits < 32 && (V - 1) >> bits) bits++;
  const int passes = bits == 0 ? 1 : (bits + 7) / 8;
  const uint32_t *pairs_dev = pairs;
  if (E && !pairs_on_device) {
    NB_TRY(cudaMalloc(&t.pairs, E * 8));
    NB_TRY(cudaMemcpyAsync(t.pairs, pairs, E * 8, cudaMemcpyHostToDevice, st));
    pairs_dev = t.pairs;
  }
  NB_TRY(cudaMalloc(&t.keys[0], En * 4));
  NB_TRY(cudaMalloc(&t.keys[1], En * 4));
  NB_TRY(cudaMalloc(&t.vals, En * 4));       // the other value buffer of the ping-pong is row_idx itself
  NB_TRY(cudaMalloc(&t.tile_hist, n_hist * 4));
  NB_TRY(cudaMalloc(&t.small, 64));          // [0] bad ids, [1] max in-degree
  NB_TRY(cudaMemsetAsync(t.small, 0, 64, st));
  NB_TRY(cudaMalloc(&t.tile_state, (size_t)scan_tiles * 8));
  NB_TRY(cudaMemsetAsync(t.tile_state, 0, (size_t)scan_tiles * 8, st));
  // one BatchParams per scan launch: k_scan tags its tile states with params->epoch, so distinct epochs need no reset in between
  BatchParams hp[8];
  memset(hp, 0, sizeof(hp));
  for (int i = 0; i < 8; i++) hp[i].epoch = (uint32_t)(i + 1);
  NB_TRY(cudaMalloc(&t.params, sizeof(hp)));
  NB_TRY(cudaMemcpyAsync(t.params, hp, sizeof(hp), cudaMemcpyHostToDevice, st));

  // values ping-pong between t.vals and g->row_idx such that the last pass lands in row_idx
  uint32_t *vbuf[2];
  vbuf[passes & 1] = g->row_idx;          // buffer index after `passes` flips (start at 0) must be row_idx
  vbuf[(passes & 1) ^ 1] = t.vals;
  if (E) {
    k_split_pairs<<<nb_grid(E, 256, 8), 256, 0, st>>>(pairs_dev, E, V, t.keys[0], vbuf[0], g->in_deg, g->out_deg, t.small);
    NB_LAUNCH_CHECK(ctx);
    for (int p = 0; p < passes; p++) {
      const int shift = 8 * p, a = p & 1, b = a ^ 1;
      k_rs_hist<<<nb_grid(n_tiles, 1, 8), RS_THREADS, 0, st>>>(t.keys[a], E, shift, n_tiles, t.tile_hist);
      NB_LAUNCH_CHECK(ctx);
      FlatScanOp op{t.tile_hist, t.tile_hist, nullptr, (uint32_t)n_hist};
      ScanWs ws = nb_scan_ws(t.tile_state, scan_tiles, t.params + p);
      k_scan<FlatScanOp><<<nb_grid(n_hist, SCAN_TILE, 4), SCAN_THREADS, 0, st>>>(op, ws);
      NB_LAUNCH_CHECK(ctx);
      k_rs_scatter<<<nb_grid(n_tiles, 1, 4), RS_THREADS, 0, st>>>(t.keys[a], vbuf[a], E, shift, n_tiles, t.tile_hist, t.keys[b], vbuf[b]);
      NB_LAUNCH_CHECK(ctx);
    }
  }
  {  // column_offset = exclusive scan of the in-degree histogram (before the clamp)
    FlatScanOp op{g->in_deg, g->col_off, g->col_off + V, V};
    ScanWs ws = nb_scan_ws(t.tile_state, scan_tiles, t.params + 6);
    k_scan<FlatScanOp><<<nb_grid(V, SCAN_TILE, 4), SCAN_THREADS, 0, st>>>(op, ws);
    NB_LAUNCH_CHECK(ctx);
  }
  k_finish_degrees<<<nb_grid(V, 256, 8), 256, 0, st>>>(g->in_deg, g->out_deg, V, t.small + 1);
  NB_LAUNCH_CHECK(ctx);
  uint32_t small_host[2] = {0, 0};
  NB_TRY(cudaMemcpyAsync(small_host, t.small, 8, cudaMemcpyDeviceToHost, st));
  NB_TRY(cudaStreamSynchronize(st));
#undef NB_TRY
  if (small_host[0]) {
    nb_set_error("nb_graph_create_from_pairs: %u edges name a vertex id >= |V| = %u", small_host[0], V);
    return fail(NB_ERR_ARG);
  }
  g->max_in_degree = small_host[1];
  *out = g;
  return NB_OK;
}
