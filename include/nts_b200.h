/* include/nts_b200.h -- C ABI of libnts_b200.so, the B200 (sm_100a) implementation of
 * NeutronOrch's sample-based training hot path.
 *
 * This is the drop-in boundary. The reference's boundary is the C++ header cuda/ntsCUDA.hpp
 * (free functions :30-71, class Cuda_Stream :177-595) implemented by cuda/ntsCUDAGraphOP.cu;
 * every entry point below names the reference interface it replaces (paths relative to the
 * reference repository). A header-only C++ adaptor with the reference's exact class and
 * method names (sample-based-gnn_b200/host/ntsCUDA.hpp) forwards to these functions, so
 * core/ and toolkits/ of the reference compile against it unchanged; INTEGRATION.md shows how.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ or torch types cross this boundary;
 *   - every function returns 0 on success, a negative nb_status otherwise, and never exits
 *     the process (the reference's wrappers print and exit(1), cuda/ntsCUDAGraphOP.cu:21-60;
 *     the adaptor keeps that convention on top of the status codes);
 *   - all work is ordered on the ctx's CUDA stream; nothing synchronises unless documented;
 *   - there is no CPU fallback: without a CUDA device every compute call fails with NB_ERR_CUDA.
 *   - VertexId = uint32_t, ValueType = float (dep/gemini/type.hpp:29-31), labels int64_t.
 */
#ifndef NTS_B200_H
#define NTS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NB_ABI_VERSION 2

typedef enum nb_status {
  NB_OK = 0,
  NB_ERR_CUDA = -1,      /* a CUDA runtime call failed; nb_last_error() has the string      */
  NB_ERR_ARG = -2,       /* invalid argument                                                 */
  NB_ERR_CAPACITY = -3,  /* a sampler arena would overflow (reference: assert, FullyRepGraph.hpp:264) */
  NB_ERR_ALIGN = -4,     /* pointer / width not usable by any vector path                    */
  NB_ERR_NCCL = -5,
  NB_ERR_UNSUPPORTED = -6
} nb_status;

/* weight applied to a sampled edge (src,dst). core/ntsFastSampler.hpp:27 (enum WeightType)
 * and :1111-1119; NB_WEIGHT_MEAN_SAMPLED is the GPU kernel's variant
 * (cuda/ntsCUDATransferKernel.cuh:319-342: divide by the sampled column length). */
typedef enum nb_weight_type {
  NB_WEIGHT_SUM = 0,           /* 1/(sqrt(out_deg[src])*sqrt(in_deg[dst]))                  */
  NB_WEIGHT_MEAN = 1,          /* the above / in_deg[dst]                                   */
  NB_WEIGHT_NONE = 2,          /* no weights computed                                       */
  NB_WEIGHT_MEAN_SAMPLED = 3   /* the above / (number of sampled in-edges of dst)           */
} nb_weight_type;

/* sampler flags */
#define NB_SAMPLER_MERGE_SRC_DST 1u /* every dst is also a src; dst_local_id is produced
                                       (sampCSC::set_merge_src_dst, core/coocsc.hpp:407-412)      */
#define NB_SAMPLER_UP_DEGREE     2u /* weights use per-batch sampled degrees (cfg UP_DEGREE:1;
                                       core/FullyRepGraph.hpp:189-226)                            */
#define NB_SAMPLER_BUILD_CSR     4u /* also build row_offset / column_indices / e_w_b
                                       (sampCSC::csc_to_csr, core/coocsc.hpp:82-111)             */

#define NB_SAMPLER_NO_BOTTOM_CSR  8u /* with BUILD_CSR: skip the CSR of the bottom layer. Its backward runs only when the
                                       bottom op's input needs a gradient, which it never does when that input is the gathered
                                       feature leaf (GCN/GraphSAGE toolkits; core/ntsContext.hpp:443 stops before the first op) */

typedef struct nb_ctx nb_ctx;         /* replaces class Cuda_Stream, cuda/ntsCUDA.hpp:177-199     */
typedef struct nb_graph nb_graph;     /* replaces the device side of FullyRepGraph +
                                         FastSampler's GPU ctor, core/ntsFastSampler.hpp:125-176  */
typedef struct nb_sampler nb_sampler; /* replaces SampledSubgraph's GPU arenas and per-layer
                                         state, core/FullyRepGraph.hpp:91-131, 258-524            */
typedef struct nb_table nb_table;     /* HBM-resident (optionally GPU-sharded) feature table:
                                         replaces GNNDatum's pinned table + per-GPU cache,
                                         core/ntsDataloador.hpp:187,483; GS_SAMPLE_PC_MULTI.hpp:916-1015 */

/* One sampled layer, all pointers in device memory and owned by the sampler; valid until the
 * next nb_sampler_sample()/replay() on the same sampler. Mirrors class sampCSC's dev_* members
 * (core/coocsc.hpp:427-461). */
typedef struct nb_layer_view {
  uint32_t n_dst, n_edges, n_src, reserved;
  const uint32_t *destination;    /* [n_dst]   global ids            (dev_destination)          */
  const uint32_t *column_offset;  /* [n_dst+1]                       (dev_column_offset)        */
  const uint32_t *sample_ans;     /* [n_edges] global src per edge   (host-only in reference)   */
  const uint32_t *row_indices;    /* [n_edges] local src ids         (dev_row_indices)          */
  const uint32_t *source;         /* [n_src]   global ids ASCENDING  (dev_source)               */
  const uint32_t *row_offset;     /* [n_src+1] CSR                   (dev_row_offset)           */
  const uint32_t *column_indices; /* [n_edges] local dst ids, CSR    (dev_column_indices)       */
  const uint32_t *csr_to_csc;     /* [n_edges] CSC position of each CSR entry (new)             */
  const float *edge_weight_forward;  /* [n_edges] CSC order (dev_edge_weight_forward / edge_weight) */
  const float *edge_weight_backward; /* [n_edges] CSR order (dev_edge_weight_backward)          */
  const uint32_t *dst_local_id;   /* [n_dst] row of each dst inside source (dev_dst_local_id)   */
  const uint32_t *src_to_dst;     /* [n_src] index of the dst equal to this src, or 0xffffffff (new) */
  const uint32_t *source_use_count; /* [n_src] number of this layer's edges that read each source (new; the CSR row lengths) */
  const uint32_t *gather_index;   /* [n_edges], bottom layer only (NULL elsewhere, and when |V| >= 2^31): sample_ans[e] with bit 31
                                     set when the batch reads that source at least "gather_keep_min_uses" (3) times -- the index + L2 hint that
                                     nb_aggregate_gathered_fwd_dyn consumes (new) */
} nb_layer_view;

/* ---- library ---------------------------------------------------------------------------- */
int nb_abi_version(void);
const char *nb_last_error(void); /* thread-local, valid until the next failing call */
int nb_device_count(int *count);
/* tuning knobs (also readable from the environment at first use):
 *   "gather_variant" / NB_GATHER_VARIANT : 0 = register path (LDG/STG), 1 = TMA bulk copies when rows are 16-byte aligned (default),
 *       2 = nb_gather_rows through TMA tensor maps, four rows per instruction (cp.async.bulk.tensor.2d tile::gather4 + tiled tensor stores),
 *       falling back where a shape is not eligible. Measured (profiles/r2b_gather4_ab.txt): F=602 117.6 us vs 120.7 (variant 1); F=100 / 128
 *       28.7 / 27.9 us vs 22.5 / 24.5 for the register path with 4 rows per warp -- not the default for any shape
 *   "table_gather_tma" : 0 (default) = nb_table_gather reads shard rows through the register path; 1 = as TMA bulk copies. Measured
 *       identical over NVLink (profiles/r2b_shard_gather_tma_ab_n2.txt)
 *   "mirror_host_tables" / NB_MIRROR_HOST_TABLES : 1 = a feature table found in mapped pinned HOST memory (the reference's zero-copy
 *       table, core/ntsDataloador.hpp:187) is copied to HBM once, on the first gather that sees it, and read from HBM afterwards
 *       (the buffer must not change after that); 0 (default) = gather over PCIe like the reference
 *   "mirror_host_adjacency" / NB_MIRROR_HOST_ADJACENCY : 1 (default) = the same for the adjacency array that the stage-shaped
 *       sampling calls receive as a mapped host pointer (core/ntsFastSampler.hpp:159-166); the topology never changes after load
 *   "sampler_fused" / NB_SAMPLER_FUSED : 0 (default) = samplers created afterwards use the general kernels (single-pass look-back scans
 *       in global memory, 256-thread blocks of <= 40 registers: they fit beside the running aggregation's blocks); 1 = the small-shape
 *       kernels (a layer's prefix sums in shared memory, fewer launches) wherever a layer fits. Identical results; measured on the
 *       Reddit shape the general kernels are faster alone (80 vs 110-135 us per batch) and beside the aggregation
 *       (profiles/r2_sweep_sampler_residency*.txt). "sampler_block_threads" (256 | 512) shapes the small-shape path.
 *   "sampler_csr_branch" : 1 (default) = in the captured batch graph a layer's CSR kernels run on a parallel branch beside the next layer's
 *       sampling (90 vs 103 us per batch alone); "sampler_tail" : bit 0 / bit 1 = the popcount scan of the dedup bitmap / the next layer's
 *       count scan are left to the last block of the sampling / relabel kernel instead of a scan kernel (default 0: a single block's scan of
 *       25K items loses to the look-back scan kernel, profiles/r2_sampler_tail_ab.txt). Identical results either way.
 *   "sampler_blocks_per_sm" : 2 (default) = no kernel of a sampler's batch graph launches more than this many blocks per SM, so that a
 *       batch sampled beside the previous batch's aggregation lives in the slot that kernel leaves free instead of displacing its
 *       blocks (0.146 -> 0.139 ms per step); 0 = every kernel sized for its own work (lowest latency on an idle GPU: 80 vs 86 us)
 *   "sampler_two_level" : -1 (default) = samplers created afterwards pick the dedup bitmap layout by density (two levels when the graph
 *       has several times more bitmap words than a batch can touch: O(|V|/1024 + S + E) per layer instead of O(|V|/32)); 0 / 1 force it
 *   "gather_keep_min_uses" : 3 (default) = a source row that the batch's bottom layer reads at least this many times gets the
 *       "keep in L2" hint bit in nb_layer_view.gather_index (read with evict_last by nb_aggregate_gathered_fwd_dyn); other rows stream
 *       through (evict_first). Changes cache behaviour only, never results.
 *   "trace" / NB_TRACE : 1 = wall-clock time spent inside every entry point is accumulated (host side); 2 = the call's stream is
 *       synchronised before the clock stops (host + GPU time per call; serialises, diagnostic only). The table goes to stderr at
 *       exit or through nb_trace_dump(). Replaces the reference's get_time() accumulators (core/ntsFastSampler.hpp:30-37) and
 *       Cuda_Stream::cpu_inclusiveTime / inclusiveTime (cuda/ntsCUDA.hpp:180-198).
 *   A mirror is keyed by the host allocation it copies and lives until that allocation is released with nb_free_host (ntsFreeHost)
 *   or nb_mirror_invalidate(ptr) is called; a lookup that finds another allocation (a different size) at the same base drops the
 *   stale copy and mirrors again. The mirrored buffer must not be WRITTEN while its mirror is alive (call nb_mirror_invalidate
 *   after changing it); memory released with cudaFreeHost directly, bypassing this library, must be invalidated by the caller. */
int nb_set_option(const char *name, int value);
int nb_mirror_invalidate(const void *host_ptr);
int nb_trace_dump(void);
int nb_trace_reset(void);

/* ---- context: class Cuda_Stream (cuda/ntsCUDA.hpp:177-199; cuda/ntsCUDAGraphOP.cu:203-262) ----
 * nb_ctx_create      <- Cuda_Stream::Cuda_Stream()     (adopt_stream == 0: creates a non-blocking stream;
 *                       otherwise runs on `cuda_stream` as given -- NULL is then the legacy default
 *                       stream -- which is how the toolkits wrap torch's pool streams,
 *                       toolkits/GCN_SAMPLE_GPU.hpp:444-466)
 * nb_ctx_set_stream  <- Cuda_Stream::setNewStream()    (adopts a caller stream; unlike the
 *                       reference the previous own stream is destroyed only if we made it)
 * nb_ctx_stream      <- Cuda_Stream::getStream()
 * nb_ctx_synchronize <- Cuda_Stream::CUDA_DEVICE_SYNCHRONIZE() (stream synchronise)
 * nb_ctx_destroy     <- Cuda_Stream::destory_Stream() */
int nb_ctx_create(int device, void *cuda_stream, int adopt_stream, nb_ctx **out);
int nb_ctx_destroy(nb_ctx *ctx);
int nb_ctx_set_stream(nb_ctx *ctx, void *cuda_stream);
void *nb_ctx_stream(nb_ctx *ctx);
int nb_ctx_device(nb_ctx *ctx);
int nb_ctx_synchronize(nb_ctx *ctx);
uint64_t nb_ctx_launch_count(nb_ctx *ctx); /* kernels this ctx has launched (bench "gpu_launches") */

/* ---- memory helpers: free functions cuda/ntsCUDA.hpp:30-71 -------------------------------
 * nb_malloc_pinned <- cudaMallocPinned (mapped, portable)   nb_free_host    <- ntsFreeHost
 * nb_device_pointer<- getDevicePointer                       nb_malloc_device<- cudaMallocGPU / allocate_gpu_buffer/edge
 * nb_free_device   <- FreeBuffer / FreeEdge                  nb_memcpy_*     <- move_bytes_in/out(_async)
 * byte counts are size_t (the reference's `int size` helpers overflow beyond 2 GiB). */
int nb_malloc_pinned(size_t bytes, void **out);
int nb_free_host(void *p);
int nb_device_pointer(void *host_mapped, void **out);
int nb_malloc_device(size_t bytes, void **out);
int nb_free_device(void *p);
int nb_memcpy_h2d(nb_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes, int sync);
int nb_memcpy_d2h(nb_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes, int sync);
int nb_memset_async(nb_ctx *ctx, void *dst_dev, int value, size_t bytes); /* cudaSetMemAsync */

/* ---- global graph ------------------------------------------------------------------------
 * Replaces: FullyRepGraph::column_offset / row_indices (core/FullyRepGraph.hpp:700-701,724-798)
 * as FastSampler's GPU ctor ships them (core/ntsFastSampler.hpp:156-166: offsets to device,
 * adjacency left in mapped host memory and read over PCIe) and move_degree_to_gpu
 * (cuda/ntsCUDA.hpp:492-493). Here the whole in-edge CSC lives in HBM.
 * in_degree/out_degree are Graph::in/out_degree_for_backward (clamped >= 1, core/graph.hpp:4525-4530);
 * either may be NULL, then they are derived on the device from the CSC with the same clamp. */
int nb_graph_create(nb_ctx *ctx, uint32_t n_vertices, uint64_t n_edges, const uint32_t *column_offset_host,
                    const uint32_t *row_indices_host, const uint32_t *in_degree_host, const uint32_t *out_degree_host,
                    nb_graph **out);
/* The same from the raw edge list (EDGE_FILE format: little-endian (u32 src, u32 dst) pairs, core/FullyRepGraph.hpp:738-795),
 * built ON THE DEVICE: replaces FullyRepGraph::GenerateAll's host counting sort (core/FullyRepGraph.hpp:724-798). Column = dst,
 * entries of a column in file order (a stable LSD radix sort by dst, hand-written: csrc/ingest.cu), degrees clamped >= 1.
 * `pairs` is host memory (pageable or pinned) or, with pairs_on_device, 8-byte aligned device memory; ids >= n_vertices are
 * rejected (NB_ERR_ARG). Bit-identical to nb_graph_create on the host-built arrays. */
int nb_graph_create_from_pairs(nb_ctx *ctx, uint32_t n_vertices, uint64_t n_edges, const uint32_t *pairs, int pairs_on_device,
                               nb_graph **out);
/* The same from CSC arrays that already live in device memory (generated or loaded on the GPU); degrees are derived on the device. */
int nb_graph_create_from_device(nb_ctx *ctx, uint32_t n_vertices, uint64_t n_edges, const uint32_t *column_offset_dev,
                                const uint32_t *row_indices_dev, nb_graph **out);
int nb_graph_destroy(nb_graph *g);
int nb_graph_info(nb_graph *g, uint32_t *n_vertices, uint64_t *n_edges, const uint32_t **column_offset_dev,
                  const uint32_t **row_indices_dev, const uint32_t **in_degree_dev, const uint32_t **out_degree_dev);

/* ---- feature / label / mask files (host side) ------------------------------------------------------
 * nb_read_feature_table <- GNNDatum::readFeature_Label_Mask's feature part (core/ntsDataloador.hpp:999-1063): the text file
 *                          ("id v0 ... v{F-1}" per line, vertices in any order) parsed by all host threads with strtof (the conversion
 *                          operator>> ends in: bit-identical values) into out[(id - id_begin) * F ...] for id in [id_begin, id_end).
 *                          use_binary_cache: a raw copy (<path>.nb_f32: "NBF1", |V|, F, floats) is written after the first parse and
 *                          read instead of the text on later calls while it is at least as new as the text file.
 * nb_read_label_mask    <- the label ("id label") and mask ("id train|eval|val|test") parts of the same reader. */
int nb_read_feature_table(const char *path, uint32_t n_vertices, uint32_t feature_size, uint32_t id_begin, uint32_t id_end, float *out,
                          int use_binary_cache, int *from_cache_out);
int nb_read_label_mask(const char *label_path, const char *mask_path, uint32_t id_begin, uint32_t id_end, int64_t *label_out, int32_t *mask_out);

/* ---- sampler -----------------------------------------------------------------------------
 * nb_sampler_create <- SampledSubgraph(layers, batch, fanout, |V|, Cuda_Stream*) (FullyRepGraph.hpp:91-131).
 *   Arenas are sized from max_batch * prod(fanout) (bounded by |V| and |E|) and checked, never
 *   asserted. fanout[i] == -1 means take every in-neighbour (CPU semantics, ntsFastSampler.hpp:1003-1006);
 *   max_edges_hint bounds such a layer (0 = derive from the maximum in-degree).
 * nb_sampler_sample <- FastSampler::sample_gpu_fast / sample_gpu_fast_omit (ntsFastSampler.hpp:648-915)
 *   and, stage by stage, Cuda_Stream::sample_processing_get_co_gpu[_omit],
 *   sample_processing_traverse_gpu, set_dst_local_index, sample_processing_update_ri_gpu,
 *   ReFreshDegree/UpdateDegree[Cache], GetWeight/GetMeanWeight (cuda/ntsCUDA.hpp:331-368,420-437,506-519,562).
 *   One call samples every layer of a mini-batch with no host round trip; the only device->host
 *   traffic is the final copy of the per-layer sizes. Sampling is uniform without replacement
 *   per dst (all in-neighbours, in stored order, when deg <= fanout), driven by Philox4x32-10
 *   keyed by (rng_seed, rng_offset, layer, dst slot): reproducible, unlike the reference's
 *   random_device-seeded LCG (cuda/ntsCUDAGraphOP.cu:1607-1608).
 *   Local ids follow the CPU sampler's order -- `source` ascending by global id
 *   (ntsFastSampler.hpp:1062-1083) -- not the reference GPU path's atomic arrival order.
 *   omit_flag (device, [|V|], may be NULL): bottom-layer dst v gets 0 edges when
 *     omit_value != 0xffffffff ? omit_flag[v] == omit_value : omit_flag[v] != 0xffffffff
 *   (cuda/ntsCUDATransferKernel.cuh:771-822).
 *   views_out (host, [n_layers], may be NULL) receives each layer's view when `sync` is non-zero
 *   (one stream synchronise at the end of the batch). With sync == 0 nothing waits: call
 *   nb_ctx_synchronize() and then nb_sampler_layer() before reading sizes on the host; the
 *   device pointers are fixed per sampler, so dependent kernels can be enqueued without waiting.
 * nb_sampler_replay: same pipeline but the neighbour draws are supplied: sample_ans_host[i] is
 *   layer i's global src id per edge in column order (what sampCSC::sample_ans holds). This is
 *   the bit-exact replay path. n_edges_host[i] must equal the column-offset total.
 * nb_sampler_layer  <- SampledSubgraph::sampled_sgs[i] (device members). */
int nb_sampler_create(nb_ctx *ctx, nb_graph *g, int n_layers, const int *fanout, uint32_t max_batch, uint32_t flags,
                      uint64_t max_edges_hint, nb_sampler **out);
int nb_sampler_destroy(nb_sampler *s);
int nb_sampler_sample(nb_sampler *s, const uint32_t *seeds, uint32_t n_seeds, int seeds_on_device, uint64_t rng_seed,
                      uint64_t rng_offset, int weight_type, const uint32_t *omit_flag_dev, uint32_t omit_value,
                      nb_layer_view *views_out, int sync);
int nb_sampler_replay(nb_sampler *s, const uint32_t *seeds_host, uint32_t n_seeds, const uint32_t *const *sample_ans_host,
                      const uint32_t *n_edges_host, int weight_type, nb_layer_view *views_out);
int nb_sampler_wait(nb_sampler *s, nb_layer_view *views_out); /* completes a sync == 0 nb_sampler_sample (event wait on that batch only) */
int nb_sampler_layer(nb_sampler *s, int layer, nb_layer_view *out);
/* Device addresses of layer sizes (and the arena capacities that bound them), for the *_dyn entry points below:
 * kernels that consume a sampled layer read their extents from device memory, so sampling, gather and
 * aggregation of a mini-batch can be enqueued back to back with no host synchronisation in between
 * (the reference synchronises after get_co, after traverse and after every gather). */
int nb_sampler_sizes_dev(nb_sampler *s, int layer, const uint32_t **n_dst_dev, const uint32_t **n_edges_dev,
                         const uint32_t **n_src_dev, uint32_t *cap_dst, uint32_t *cap_edges, uint32_t *cap_src);

/* ---- stage-by-stage sampling, in the reference's own call shapes ------------------------------------------------
 * For callers that drive the stages themselves with host round trips in between (SampledSubgraph::gpu_sampling_init_co,
 * gpu_sampling, update_degrees_GPU, Get_Weight -- core/FullyRepGraph.hpp:213-239, 326-524). Same kernels as
 * nb_sampler_sample, on caller-owned arrays; transient state lives in the ctx scratch buffer.
 * nb_sample_count        <- Cuda_Stream::sample_processing_get_co_gpu / _omit (cuda/ntsCUDA.hpp:331-349, 514-519); synchronises
 *                           and returns the edge count, as the reference does through its VertexId_CUDA& edge_size.
 * nb_sample_traverse     <- Cuda_Stream::sample_processing_traverse_gpu (:355-368): r_i receives GLOBAL src ids per edge;
 *                           src receives the distinct ids ASCENDING (reference: atomic arrival order); src_index[v] = local id
 *                           of every sampled v (reference: the same |V| map); *src_count_dev = number of distinct ids.
 * nb_sample_update_ri    <- Cuda_Stream::sample_processing_update_ri_gpu (:351-354): r_i[e] = src_index[r_i[e]].
 * nb_set_dst_local_index <- Cuda_Stream::set_dst_local_index (:562-563).
 * nb_update_degree       <- Cuda_Stream::ReFreshDegree + UpdateDegree / UpdateDegreeCache (:420-429, 506-508).
 * nb_edge_weight         <- Cuda_Stream::GetWeight / GetMeanWeight (:430-437, 510-512). */
int nb_sample_count(nb_ctx *ctx, const uint32_t *dst_dev, uint32_t *local_column_offset_dev, const uint32_t *global_column_offset_dev,
                    uint32_t dst_size, uint32_t fanout, const uint32_t *omit_flag_dev, uint32_t omit_value, uint32_t *edge_size_out);
int nb_sample_traverse(nb_ctx *ctx, const uint32_t *destination_dev, const uint32_t *column_offset_dev, uint32_t *r_i_dev,
                       const uint32_t *global_column_offset_dev, const uint32_t *global_row_indices_dev, uint32_t *src_index_dev,
                       uint32_t vtx_size, uint32_t edge_size, uint32_t n_vertices, uint32_t *src_dev, uint32_t *src_count_dev,
                       uint32_t layer, uint32_t fanout, int add_dst_to_src, uint64_t rng_seed, uint64_t rng_offset);
int nb_sample_update_ri(nb_ctx *ctx, uint32_t *r_i_dev, const uint32_t *src_index_dev, uint32_t edge_size);
int nb_set_dst_local_index(nb_ctx *ctx, const uint32_t *src_index_dev, const uint32_t *destination_dev, uint32_t n_dst,
                           uint32_t *dst_to_local_dev);
int nb_update_degree(nb_ctx *ctx, uint32_t *out_degree_dev, uint32_t *in_degree_dev, uint32_t n_vertices, uint32_t n_dst,
                     const uint32_t *destination_dev, const uint32_t *source_dev, const uint32_t *column_offset_dev,
                     const uint32_t *row_indices_dev, int cache_fanout);
int nb_edge_weight(nb_ctx *ctx, float *edge_weight_dev, const uint32_t *out_degree_dev, const uint32_t *in_degree_dev, uint32_t n_dst,
                   const uint32_t *destination_dev, const uint32_t *source_dev, const uint32_t *column_offset_dev,
                   const uint32_t *row_indices_dev, int mean);

/* ---- hotness-aware cache selection ------------------------------------------------------------------------------
 * nb_hotness         <- nts::op::get_most_neighbor (core/ntsBaseOp.hpp:333-399; per super-batch, CPU in the reference): counts start
 *                       as the indicator of the super-batch seeds, are pushed `layers-1` hops along the in-edges of the full graph,
 *                       cache_num = (u32)((#non-zero counts + 1) * cache_rate) (or fixed_cache_num when != 0xffffffff, the
 *                       overload :266-330), pivot = count at descending rank cache_num, output = the first cache_num vertex ids
 *                       (ASCENDING, the serial order of the reference loop) whose count >= pivot. Synchronises to return cache_num.
 * nb_set_cache_index <- GNNDatum::set_cache_index (core/ntsDataloador.hpp:440-478) on device arrays:
 *                       cache_map[id] = super_batch_id, cache_location[id] = position in the list. */
int nb_hotness(nb_ctx *ctx, nb_graph *g, const uint32_t *seeds, uint32_t n_seeds, int seeds_on_device, int layers, float cache_rate,
               uint32_t fixed_cache_num, uint32_t *cache_ids_dev, uint32_t capacity, uint32_t *cache_num_out, uint32_t *counts_dev_or_null);
int nb_set_cache_index(nb_ctx *ctx, uint32_t *cache_map_dev, uint32_t *cache_location_dev, uint32_t super_batch_id,
                       const uint32_t *cache_ids_dev, uint32_t n);

/* ---- feature / label gather ----------------------------------------------------------------
 * nb_gather_rows        <- Cuda_Stream::zero_copy_feature_move_gpu (cuda/ntsCUDA.hpp:370-374): out[i,:] = table[ids[i],:].
 *                          `table` may be device memory or mapped pinned host memory; table_pitch
 *                          is the row pitch in floats (>= feature_size; the reference is always dense).
 *                          A pitch larger than feature_size declares ROW PADDING (here and in the *_dyn aggregation calls):
 *                          columns [feature_size, min(pitch, feature_size rounded up to 8)) of every output row may be
 *                          written (pad columns of the input flow into pad columns of the output), so that rows move as
 *                          128-bit vectors / whole 32-byte sectors. Dense tensors (the reference's only layout) pass
 *                          pitch == feature_size; a column slice of a wider tensor whose neighbouring columns hold data must
 *                          be made contiguous by the caller first.
 * nb_gather_rows_cached <- FastSampler::load_feature_gpu_cache (core/ntsFastSampler.hpp:263-317) =
 *                          zero_copy_feature_move_gpu_cache + gather_feature_from_gpu_cache (:378-383) with
 *                          the hot/cold split done on the device in the same kernel:
 *                          slot = cache_node_hashmap[ids[i]]; slot != -1 ? cache_table[slot,:] : cold_table[ids[i],:].
 * nb_gather_rows_indexed<- Cuda_Stream::zero_copy_feature_move_gpu_cache (slot_map == NULL) and
 *                          Cuda_Stream::gather_feature_from_gpu_cache (slot_map = cache_node_hashmap) (cuda/ntsCUDA.hpp:378-383;
 *                          kernels cuda/ntsCUDATransferKernel.cuh:154-183), the pair FastSampler::load_feature_gpu_cache issues after
 *                          its CPU hot/cold split (core/ntsFastSampler.hpp:284-312): for i in [0, n):
 *                              lid = local_idx[i]; v = ids[lid]; out[lid,:] = table[slot_map ? slot_map[v] : v, :]
 *                          local_idx and slot_map may be mapped pinned host memory (the toolkits fill both on the CPU,
 *                          GS_SAMPLE_PC_MULTI.hpp:955-1013). Asynchronous; the adaptor synchronises like the reference does, because
 *                          the caller rewrites local_idx for the next batch.
 * nb_gather_labels      <- Cuda_Stream::global_copy_label_move_gpu (:384-387).
 * nb_row_override       <- Cuda_Stream::dev_load_share_embedding (:534-537), dev_load_share_aggregate (:495-497):
 *                          rows i with cache_map[destination[i]] == super_batch_id (or != -1 when
 *                          super_batch_id == 0xffffffff) become share[cache_location[destination[i]],:].
 * nb_row_override2      <- Cuda_Stream::dev_load_share_embedding_and_feature (:521-525): two tensors at once. */
int nb_gather_rows(nb_ctx *ctx, float *out, const float *table, const uint32_t *ids_dev, uint32_t n_rows,
                   uint32_t feature_size, uint32_t table_pitch, uint32_t out_pitch);
int nb_gather_rows_dyn(nb_ctx *ctx, float *out, const float *table, const uint32_t *ids_dev, const uint32_t *n_rows_dev,
                       uint32_t max_rows, uint32_t feature_size, uint32_t table_pitch, uint32_t out_pitch);
int nb_gather_rows_cached(nb_ctx *ctx, float *out, const float *cold_table, uint32_t cold_pitch, const float *cache_table,
                          uint32_t cache_pitch, const uint32_t *cache_node_hashmap_dev, const uint32_t *ids_dev,
                          uint32_t n_rows, uint32_t feature_size, uint32_t out_pitch, uint32_t *hit_count_dev_or_null);
int nb_gather_rows_indexed(nb_ctx *ctx, float *out, uint32_t out_pitch, const float *table, uint32_t table_pitch,
                           const uint32_t *ids_dev, const uint32_t *local_idx, const uint32_t *slot_map_or_null, uint32_t n,
                           uint32_t feature_size);
int nb_gather_labels(nb_ctx *ctx, int64_t *out, const int64_t *labels_dev, const uint32_t *ids_dev, uint32_t n);
int nb_row_override(nb_ctx *ctx, float *out, const float *share, const uint32_t *cache_map_dev,
                    const uint32_t *cache_location_dev, const uint32_t *destination_dev, uint32_t n_dst,
                    uint32_t feature_size, uint32_t super_batch_id);
int nb_row_override2(nb_ctx *ctx, float *out_feature, float *out_embedding, const float *share_feature,
                     const float *share_embedding, const uint32_t *cache_map_dev, const uint32_t *cache_location_dev,
                     const uint32_t *destination_dev, uint32_t n_dst, uint32_t feature_size, uint32_t embedding_size,
                     uint32_t super_batch_id);

/* Cold-row staging (feature table too large for, or simply left in, host memory; hot rows in an HBM cache table).
 * Replaces FastSampler::load_feature_gpu_cache's CPU split + zero-copy reads (core/ntsFastSampler.hpp:263-317):
 * nb_stage_submit splits the id list on the device, ships only the cold ids to the host and returns; a worker thread packs
 * those rows into pinned memory and issues ONE cudaMemcpyAsync on a side stream; nb_stage_gather orders the merge kernel
 * (hot rows from the cache table, cold rows from the staged block) behind that copy. Two slots: submit batch i+1 while batch i
 * trains. host_table may be any host memory (pinned or pageable); it is read by the worker thread only. */
typedef struct nb_stage nb_stage;
int nb_stage_create(nb_ctx *ctx, const float *host_table, uint32_t host_pitch, uint32_t feature_size, uint32_t max_rows, nb_stage **out);
int nb_stage_destroy(nb_stage *s);
int nb_stage_submit(nb_stage *s, int slot, const uint32_t *ids_dev, uint32_t n_rows, const uint32_t *cache_node_hashmap_dev);
int nb_stage_gather(nb_stage *s, int slot, float *out, uint32_t out_pitch, const float *cache_table, uint32_t cache_pitch,
                    const uint32_t *cache_node_hashmap_dev, const uint32_t *ids_dev, uint32_t *n_cold_out);
/* The same with the hot cache partitioned over the GPUs of the node (nb_table below; cache slot k lives on shard k % n at row
 * k / n and is read over NVLink inside the kernel): GS_SAMPLE_PC_MULTI's replicated cache (toolkits/GS_SAMPLE_PC_MULTI.hpp:916-1015)
 * turned into one N-times larger cache. */
int nb_stage_gather_table(nb_stage *s, int slot, float *out, uint32_t out_pitch, nb_table *hot_table,
                          const uint32_t *cache_node_hashmap_dev, const uint32_t *ids_dev, uint32_t *n_cold_out);

/* Sharded HBM feature table (multi-GPU): row v lives on shard v % n_shards at local row
 * v / n_shards. shard_ptrs[k] is a device pointer valid on THIS device (local allocation or a
 * peer mapping obtained through CUDA IPC); the gather kernel reads peers directly over NVLink.
 * Replaces the per-GPU replicated cache of GS_SAMPLE_PC_MULTI.hpp:916-1015. */
int nb_table_create(nb_ctx *ctx, uint32_t n_shards, const float *const *shard_ptrs, uint32_t feature_size,
                    uint32_t pitch, uint64_t n_rows_total, nb_table **out);
int nb_table_destroy(nb_table *t);
int nb_table_gather(nb_ctx *ctx, nb_table *t, float *out, const uint32_t *ids_dev, uint32_t n_rows, uint32_t out_pitch);
/* CUDA IPC plumbing for the above (64-byte opaque handles) */
int nb_ipc_get_handle(void *dev_ptr, void *handle64_out);
int nb_ipc_open_handle(const void *handle64, void **dev_ptr_out);
int nb_ipc_close_handle(void *dev_ptr);
/* Peer-shareable HBM for the shards (CUDA virtual memory management, 2 MB pages on both sides of the link). Mappings made by
 * cudaIpcOpenMemHandle are translated through small pages, and a random-row gather over a multi-GB remote shard becomes
 * TLB-miss bound (measured ~45 GB/s per peer against ~740 GB/s through these mappings). nb_vmm_alloc returns device memory on
 * ctx's device and, when fd_out != NULL, a POSIX file descriptor to hand to the other processes (SCM_RIGHTS / pidfd_getfd);
 * nb_vmm_import maps the allocation behind such a descriptor for ctx's device (bytes = the exporter's request; both sides
 * round up to nb_vmm_padded_size); nb_vmm_grant lets another device of THIS process read/write the mapping (the reference's
 * one-process-many-GPUs threading); nb_vmm_free unmaps and releases either kind. There is no reference counterpart: the
 * reference replicates its cache per GPU (GS_SAMPLE_PC_MULTI.hpp:916-1015). */
size_t nb_vmm_padded_size(nb_ctx *ctx, size_t bytes);
int nb_vmm_alloc(nb_ctx *ctx, size_t bytes, void **dev_ptr_out, int *fd_out);
int nb_vmm_import(nb_ctx *ctx, int fd, size_t bytes, void **dev_ptr_out);
int nb_vmm_grant(void *dev_ptr, int device);
int nb_vmm_free(void *dev_ptr);

/* ---- dense-gradient exchange over NVLink peer memory -----------------------------------------------
 * Replaces Parameter::reduce_multi_gpu_gradient -> NCCL_Communicator::AllReduce (core/NtsScheduler.hpp:830-836,
 * cuda/ntsCUDAGraphOP.cu:173-200): in-place SUM over the ranks of one node of a small fp32 buffer (the dense weight gradients,
 * ~330 KB per step) by this library's own kernels over peer-mapped memory. Every rank's block (nb_peer_comm_block_bytes,
 * allocated with nb_vmm_alloc and mapped by every other rank with nb_vmm_import) holds arrival flags and two slots of `world`
 * regions. blocks[r] = rank r's block as mapped on ctx's device (blocks[rank] = the local allocation).
 * nb_peer_allreduce_begin : PUSH -- this rank's buffer is written into its region of every rank's slot (remote stores over
 *                           NVLink), then a flag per chunk; waits for nobody. The push kernel is ordered behind everything
 *                           enqueued on ctx's stream so far but runs on a side stream of the communicator (option
 *                           "peer_push_side_stream", default 1; 0 = in ctx's stream), beside what the caller enqueues next:
 *                           `in` must stay unchanged until nb_peer_allreduce_end has been enqueued.
 * nb_peer_allreduce_end   : REDUCE -- waits (bounded, ~20 s, then nb_peer_comm_check reports it) for every rank's flags and
 *                           sums the regions of the LOCAL slot in rank order: bit-identical on every rank, deterministic.
 *                           Whatever the caller enqueues between begin and end (the next batch's gather + aggregation)
 *                           absorbs rank skew; the reduce kernel (the only one that waits) runs in ctx's stream behind the
 *                           local push, so it never waits for SM slots behind a persistent kernel of another stream.
 *                           One exchange in flight per communicator.
 * nb_peer_allreduce_sum   : both phases in ONE launch.
 * Every rank issues the same sequence of calls with the same n.
 * nb_peer_comm_stats      : exchanges ended, and the time their reduce phases spent polling for the slowest peer (sum / max, ns)
 *                           since the last reset -- the rank-skew the exchange could not hide, as a number. */
typedef struct nb_peer_comm nb_peer_comm;
size_t nb_peer_comm_block_bytes(uint64_t max_floats, uint32_t world);
int nb_peer_comm_create(nb_ctx *ctx, uint32_t rank, uint32_t world, uint64_t max_floats, void *const *blocks, nb_peer_comm **out);
int nb_peer_comm_destroy(nb_peer_comm *c);
int nb_peer_allreduce_sum(nb_peer_comm *c, float *inout, uint64_t n);
int nb_peer_allreduce_begin(nb_peer_comm *c, const float *in, uint64_t n);
int nb_peer_allreduce_end(nb_peer_comm *c, float *out, uint64_t n);
int nb_peer_comm_check(nb_peer_comm *c, int *timed_out);
int nb_peer_comm_stats(nb_peer_comm *c, uint64_t *exchanges, uint64_t *wait_ns_sum, uint64_t *wait_ns_max, int reset);

/* ---- sparse aggregation ----------------------------------------------------------------------
 * nb_aggregate_csc_fwd <- Cuda_Stream::Gather_By_Dst_From_Src_Spmm (cuSPARSE, cuda/ntsCUDAGraphOP.cu:425-587) and
 *                         Gather_By_Dst_From_Src / _Optim (cuda/ntsCUDA.hpp:223-267):
 *                         output[d,:] = sum_{e in column d} w[e] * input[row_indices[e],:], column order,
 *                         multiply then add. weight == NULL means w = 1 (with_weight=false).
 *                         Writes every output row (an empty column gives zeros).
 * nb_aggregate_csr_bwd <- Cuda_Stream::Gather_By_Src_From_Dst_Spmm (:901-1042) / Gather_By_Src_From_Dst[_Optim]:
 *                         output[s,:] = sum_{j in row s} w_b[j] * input[column_indices[j],:].
 * nb_aggregate_push_bwd<- Cuda_Stream::Push_From_Dst_To_Src_Spmm (:621-770) / Push_From_Dst_To_Src:
 *                         output[row_indices[e],:] += w[e] * input[d,:] from the CSC alone
 *                         (vectorised red.global.add; output is zeroed here first; summation order unspecified). */
int nb_aggregate_csc_fwd(nb_ctx *ctx, const float *input, float *output, const float *weight_forward,
                         const uint32_t *row_indices, const uint32_t *column_offset, uint32_t n_dst, uint32_t n_src,
                         uint32_t feature_size);
int nb_aggregate_csr_bwd(nb_ctx *ctx, const float *input, float *output, const float *weight_backward,
                         const uint32_t *row_offset, const uint32_t *column_indices, uint32_t n_src, uint32_t n_dst,
                         uint32_t feature_size);
/* same operators with explicit row pitches (in floats) and the row count read from device memory
 * (n_*_dev == NULL: the max_* argument is the row count) */
int nb_aggregate_csc_fwd_dyn(nb_ctx *ctx, const float *input, float *output, const float *weight_forward,
                             const uint32_t *row_indices, const uint32_t *column_offset, const uint32_t *n_dst_dev,
                             uint32_t max_dst, uint32_t feature_size, uint32_t input_pitch, uint32_t output_pitch);
int nb_aggregate_csr_bwd_dyn(nb_ctx *ctx, const float *input, float *output, const float *weight_backward,
                             const uint32_t *row_offset, const uint32_t *column_indices, const uint32_t *n_src_dev,
                             uint32_t max_src, uint32_t feature_size, uint32_t input_pitch, uint32_t output_pitch);
/* nb_aggregate_gathered_fwd_dyn <- FastSampler::load_feature_gpu (core/ntsFastSampler.hpp:244-261) + the bottom hop's
 *                         SingleGPU[All]SampleGraphOp::forward (core/ntsSingleGPUSampleGraphOp.hpp:209-247) as ONE kernel:
 *                         output[d,:] = sum_{e in column d} w[e] * table[gather_index[e] & 0x7fffffff,:] -- the same rows in the same
 *                         order as gathering X0 = table[source] first, hence the same bits, without writing or re-reading X0.
 *                         Bit 31 of a gather_index entry (nb_layer_view::gather_index) steers L2: rows that the batch reads more
 *                         than once are kept (evict_last), single-use rows are not (evict_first). */
int nb_aggregate_gathered_fwd_dyn(nb_ctx *ctx, const float *table, uint32_t table_pitch, const uint32_t *gather_index, float *output,
                                  const float *weight_forward, const uint32_t *column_offset, const uint32_t *n_dst_dev,
                                  uint32_t max_dst, uint32_t feature_size, uint32_t output_pitch);
int nb_aggregate_push_bwd(nb_ctx *ctx, const float *input, float *output, const float *weight, const uint32_t *row_indices,
                          const uint32_t *column_offset, uint32_t n_dst, uint32_t n_src, uint32_t feature_size);

/* ---- GAT edge ops ------------------------------------------------------------------------------
 * Legacy-shaped (one per reference kernel, so unmodified BatchGPU* ops keep working):
 * nb_scatter_src_dst_to_msg <- Cuda_Stream::Scatter_Src_Dst_to_Msg (cuda/ntsCUDA.hpp:567-569)
 * nb_gather_msg_to_src_dst  <- Cuda_Stream::Gather_Msg_To_Src_Dst (:570-573) (output zeroed here)
 * nb_edge_softmax_fwd       <- Cuda_Stream::Edge_Softmax_Forward_Norm_Block (:321-324) (feature_size 1)
 * nb_edge_softmax_bwd       <- Cuda_Stream::Edge_Softmax_Backward_Block (:326-329)
 * nb_gather_msg_to_dst      <- Cuda_Stream::Gather_Msg_to_Dst (:312-314)
 * nb_scatter_dst_to_msg     <- Cuda_Stream::Scatter_Dst_to_Msg (:308-310)
 * Fused layer (replaces the five-op chain of toolkits/GAT_SAMPLE_ALL_MULTI.hpp:383-464 between X*W and relu):
 * nb_gat_fwd: score_pre[e] = h[src(e)].att[0:F] + h[dst_local_id[d]].att[F:2F];
 *             alpha = edge_softmax(leaky_relu(score_pre, slope)); out[d,:] = sum_e alpha[e]*h[src(e),:].
 * nb_gat_bwd: gradients of the above w.r.t. h and att, deterministic (CSR gather, no float atomics on dh). */
int nb_scatter_src_dst_to_msg(nb_ctx *ctx, float *message, const float *src_feature, const uint32_t *row_indices,
                              const uint32_t *column_offset, uint32_t n_dst, uint32_t feature_size,
                              const uint32_t *dst_local_id);
int nb_gather_msg_to_src_dst(nb_ctx *ctx, float *src_grad, const float *message_grad, const uint32_t *row_indices,
                             const uint32_t *column_offset, uint32_t n_dst, uint32_t n_src, uint32_t feature_size,
                             const uint32_t *dst_local_id);
int nb_edge_softmax_fwd(nb_ctx *ctx, float *msg_output, const float *msg_input, float *msg_cached,
                        const uint32_t *column_offset, uint32_t n_dst);
int nb_edge_softmax_bwd(nb_ctx *ctx, float *msg_input_grad, const float *msg_output_grad, const float *msg_cached,
                        const uint32_t *column_offset, uint32_t n_dst);
int nb_gather_msg_to_dst(nb_ctx *ctx, float *dst_feature, const float *message, const uint32_t *column_offset,
                         uint32_t n_dst, uint32_t feature_size);
int nb_scatter_dst_to_msg(nb_ctx *ctx, float *message, const float *dst_feature, const uint32_t *column_offset,
                          uint32_t n_dst, uint32_t feature_size);
int nb_gat_fwd(nb_ctx *ctx, const float *h, const float *att, float negative_slope, const uint32_t *column_offset,
               const uint32_t *row_indices, const uint32_t *dst_local_id, uint32_t n_dst, uint32_t n_src,
               uint32_t feature_size, float *score_pre, float *alpha, float *out);
int nb_gat_bwd(nb_ctx *ctx, const float *h, const float *att, float negative_slope, const float *dout,
               const float *score_pre, const float *alpha, const uint32_t *column_offset, const uint32_t *row_indices,
               const uint32_t *dst_local_id, const uint32_t *row_offset, const uint32_t *column_indices,
               const uint32_t *csr_to_csc, const uint32_t *src_to_dst, uint32_t n_dst, uint32_t n_src, uint32_t n_edges,
               uint32_t feature_size, float *dh, float *datt);

#ifdef __cplusplus
}
#endif
#endif /* NTS_B200_H */
