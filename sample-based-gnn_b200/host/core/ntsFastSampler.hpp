/* core/ntsFastSampler.hpp -- drop-in for the reference header of the same name (AiX-im/Sample-based-GNN, core/ntsFastSampler.hpp).
 *
 * Put sample-based-gnn_b200/host before the reference tree on the include path (the same -I that selects cuda/ntsCUDA.hpp): every
 * `#include "core/ntsFastSampler.hpp"` of core/ and toolkits/ then lands here. The reference's own class is included underneath,
 * renamed, and `FastSampler` derives from it: the CPU sampler, the loaders, the timers and every public member the toolkits read
 * (ssg, ssgs, work_range, work_offset, sample_nids, layer, fanout ...) are the reference's own code and data. What changes is the GPU
 * sampling entry points
 *     sample_gpu_fast(batch[, ssg_id][, weightType])                                   core/ntsFastSampler.hpp:648-709, 944-948
 *     sample_gpu_fast_omit(batch[, ssg_id], CacheFlag[, super_batch_id][, weightType]) :711-915, 950-960
 * which in the reference drive SampledSubgraph::gpu_* stage by stage with three host round trips per layer (get_co, traverse,
 * weights; core/FullyRepGraph.hpp:326-524). Here one call = nb_sampler_sample: every layer of the mini-batch as ONE CUDA graph on the
 * pipeline slot's stream, one synchronisation at the end to learn the sizes. The results are published where the reference publishes
 * them -- the slot's SampledSubgraph: sampled_sgs[i]->{v_size, e_size, src_size, dev_destination, dev_column_offset, dev_row_indices,
 * dev_source, edge_weight, dev_dst_local_id} -- as pointers into the sampler's arena (the reference's are pointers into its own
 * arena, core/FullyRepGraph.hpp:258-315), so SingleGPUAllSampleGraphOp, BatchGPU*Op, load_feature_gpu[_cache], load_label_gpu and
 * load_share_embedding work unchanged. Layer chaining holds: sampled_sgs[i+1]->dev_destination == sampled_sgs[i]->dev_source.
 *
 * Differences a caller can observe: `dev_source` is ascending by global id (the CPU sampler's order; the reference GPU path's is the
 * arrival order of its atomics), the weightType argument is honoured (the reference's ssg_id overload drops it, :944-948), and the
 * draws come from a counter-based Philox stream (seed NB_SAMPLER_SEED, default 0x5EED0004) instead of random_device.
 * NB_LEGACY_SAMPLER=1 routes the calls to the reference's stage-by-stage code (served by the same library) for A/B runs.
 */
#ifndef NTS_B200_SHADOW_NTSFASTSAMPLER_HPP
#define NTS_B200_SHADOW_NTSFASTSAMPLER_HPP

#define FastSampler NtsReferenceFastSampler
#include_next "core/ntsFastSampler.hpp"
#undef FastSampler

#include <atomic>
#include <map>
#include <mutex>
#include <stdlib.h>

#include "nts_b200.h"

class FastSampler : public NtsReferenceFastSampler {
  struct SharedGraph { nb_graph *g; int users; };
  static std::map<std::pair<FullyRepGraph *, int>, SharedGraph> &nb_graphs() { static std::map<std::pair<FullyRepGraph *, int>, SharedGraph> m; return m; }
  static std::mutex &nb_mutex() { static std::mutex m; return m; }
  static bool nb_legacy() { static int v = -1; if (v < 0) { const char *e = getenv("NB_LEGACY_SAMPLER"); v = (e && e[0] == '1') ? 1 : 0; } return v == 1; }

  nb_graph *nb_g = nullptr;
  int nb_device = -1;
  std::vector<nb_sampler *> nb_slot;   /* one arena per pipeline slot, created on first use (merge / up_degree are known by then) */
  std::vector<uint32_t> nb_cap;        /* seeds the slot's arena was sized for */
  int nb_slots = 1;
  uint32_t nb_max_batch = 0;
  std::atomic<uint64_t> nb_counter{0};
  uint64_t nb_seed = 0x5EED0004ull;

  nb_sampler *nb_get(int slot, uint32_t n_seeds) {
    std::lock_guard<std::mutex> lock(nb_mutex());
    if ((int)nb_slot.size() < nb_slots) { nb_slot.resize(nb_slots, nullptr); nb_cap.resize(nb_slots, 0); }
    if (nb_slot[slot] && n_seeds <= nb_cap[slot]) return nb_slot[slot];
    if (nb_slot[slot]) {   /* GS_SAMPLE_CACHE samples whole super-batches through a sampler it constructed with BATCH_SIZE: the reference's
                              fixed 20M-word arena absorbs that (core/FullyRepGraph.hpp:113-116); here the slot's arena is re-sized */
      NTS_B200_CHECK(nb_sampler_destroy(nb_slot[slot]));
      nb_slot[slot] = nullptr;
    }
    nb_ctx *ctx = ssgs[slot]->cs->ctx;
    if (!nb_g) {   /* the global CSC goes to HBM once per (graph, device): train / eval / test samplers share it */
      nb_device = nb_ctx_device(ctx);
      auto key = std::make_pair(whole_graph, nb_device);
      auto it = nb_graphs().find(key);
      if (it == nb_graphs().end()) {
        nb_graph *g = nullptr;
        NTS_B200_CHECK(nb_graph_create(ctx, whole_graph->global_vertices, whole_graph->global_edges, whole_graph->column_offset,
                                       whole_graph->row_indices, graph->in_degree_for_backward, graph->out_degree_for_backward, &g));
        it = nb_graphs().insert(std::make_pair(key, SharedGraph{g, 0})).first;
      }
      it->second.users++;
      nb_g = it->second.g;
    }
    const bool merge = ssgs[slot]->sampled_sgs.size() && ssgs[slot]->sampled_sgs[0]->is_merge_src_dst;
    const uint32_t flags = (merge ? NB_SAMPLER_MERGE_SRC_DST : 0u) | (graph->config->up_degree ? NB_SAMPLER_UP_DEGREE : 0u);
    if (const char *e = getenv("NB_SAMPLER_SEED")) nb_seed = strtoull(e, nullptr, 0);
    nb_cap[slot] = std::max(nb_max_batch, n_seeds);
    NTS_B200_CHECK(nb_sampler_create(ctx, nb_g, layer, fanout.data(), nb_cap[slot], flags, 0, &nb_slot[slot]));
    return nb_slot[slot];
  }

  SampledSubgraph *nb_sample(int batch_size_, int slot, VertexId *CacheFlag, VertexId omit_value, WeightType weightType) {
    SampledSubgraph *sg = ssgs[slot];
    ssg = sg;
    assert(work_offset < work_range[1]);
    const uint32_t actual = std::min((VertexId)batch_size_, work_range[1] - work_offset);
    nb_sampler *s = nb_get(slot, actual);
    nb_layer_view views[8];
    const int w = weightType == WeightType::Sum ? NB_WEIGHT_SUM : weightType == WeightType::Mean ? NB_WEIGHT_MEAN_SAMPLED : NB_WEIGHT_NONE;
    NTS_B200_CHECK(nb_sampler_sample(s, &sample_nids[work_offset], actual, 0, nb_seed, nb_counter.fetch_add(1), w, CacheFlag, omit_value, views, 1));
    for (int i = 0; i < layer; i++) {
      sampCSC *c = sg->sampled_sgs[i];
      c->v_size = views[i].n_dst; c->e_size = views[i].n_edges; c->src_size = views[i].n_src;
      c->dev_destination = (VertexId *)views[i].destination;
      c->dev_column_offset = (VertexId *)views[i].column_offset;
      c->dev_row_indices = (VertexId *)views[i].row_indices;
      c->dev_source = (VertexId *)views[i].source;
      c->edge_weight = (ValueType *)views[i].edge_weight_forward;
      if (views[i].dst_local_id) c->dev_dst_local_id = (VertexId *)views[i].dst_local_id;
    }
    sg->curr_layer = layer - 1;
    sg->curr_dst_size = views[layer - 1].n_dst;
    work_offset += actual;
    return sg;
  }
  int nb_current_slot() const {
    for (int i = 0; i < nb_slots; i++) if (ssgs[i] == ssg) return i;
    return 0;
  }

public:
  /* CPU sampler ctor (core/ntsFastSampler.hpp:73-123): unchanged behaviour */
  FastSampler(Graph<Empty> *graph_, FullyRepGraph *whole_graph_, std::vector<VertexId> &index, int layers_, std::vector<int> fanout_,
              int batch_size, bool to_gpu_ = false, int gpu_id_ = 0, int pipeline_num = 1, Cuda_Stream *cudaStreamArray = nullptr)
      : NtsReferenceFastSampler(graph_, whole_graph_, index, layers_, fanout_, batch_size, to_gpu_, gpu_id_, pipeline_num, cudaStreamArray) {}
  /* GPU sampler ctor (:125-176) */
  FastSampler(FullyRepGraph *whole_graph_, std::vector<VertexId> &index, int layers_, int batch_size_, std::vector<int> fanout_,
              int pipeline_num = 1, Cuda_Stream *cuda_stream = nullptr)
      : NtsReferenceFastSampler(whole_graph_, index, layers_, batch_size_, fanout_, pipeline_num, cuda_stream),
        nb_slots(pipeline_num <= 1 ? 1 : pipeline_num), nb_max_batch((uint32_t)batch_size_) {}
  ~FastSampler() {
    std::lock_guard<std::mutex> lock(nb_mutex());
    for (nb_sampler *s : nb_slot) if (s) nb_sampler_destroy(s);
    if (nb_g) {
      auto it = nb_graphs().find(std::make_pair(whole_graph, nb_device));
      if (it != nb_graphs().end() && --it->second.users == 0) { nb_graph_destroy(it->second.g); nb_graphs().erase(it); }
    }
  }

  SampledSubgraph *sample_gpu_fast(int batch_size_, WeightType weightType = WeightType::Sum) {
    if (nb_legacy()) return NtsReferenceFastSampler::sample_gpu_fast(batch_size_, weightType);
    return nb_sample(batch_size_, nb_current_slot(), nullptr, 0xffffffffu, weightType);
  }
  SampledSubgraph *sample_gpu_fast(int batch_size_, int ssg_id, WeightType weightType = WeightType::Sum) {
    if (nb_legacy()) return NtsReferenceFastSampler::sample_gpu_fast(batch_size_, ssg_id, weightType);
    return nb_sample(batch_size_, ssg_id, nullptr, 0xffffffffu, weightType);
  }
  /* bottom-layer dst with CacheFlag[v] != -1 get no edges (:747-763) */
  SampledSubgraph *sample_gpu_fast_omit(int batch_size_, VertexId *CacheFlag, WeightType weightType = WeightType::Sum) {
    if (nb_legacy()) return NtsReferenceFastSampler::sample_gpu_fast_omit(batch_size_, CacheFlag, weightType);
    return nb_sample(batch_size_, nb_current_slot(), CacheFlag, 0xffffffffu, weightType);
  }
  /* bottom-layer dst with CacheFlag[v] == super_batch_id get no edges (:850-862) */
  SampledSubgraph *sample_gpu_fast_omit(int batch_size_, VertexId *CacheFlag, VertexId super_batch_id, WeightType weightType = WeightType::Sum) {
    if (nb_legacy()) return NtsReferenceFastSampler::sample_gpu_fast_omit(batch_size_, CacheFlag, super_batch_id, weightType);
    return nb_sample(batch_size_, nb_current_slot(), CacheFlag, super_batch_id, weightType);
  }
  SampledSubgraph *sample_gpu_fast_omit(int batch_size_, int ssg_id, VertexId *CacheFlag, WeightType weightType = WeightType::Sum) {
    if (nb_legacy()) return NtsReferenceFastSampler::sample_gpu_fast_omit(batch_size_, ssg_id, CacheFlag, weightType);
    return nb_sample(batch_size_, ssg_id, CacheFlag, 0xffffffffu, weightType);
  }
  SampledSubgraph *sample_gpu_fast_omit(int batch_size_, int ssg_id, VertexId *CacheFlag, VertexId super_batch_id,
                                        WeightType weightType = WeightType::Sum) {
    if (nb_legacy()) return NtsReferenceFastSampler::sample_gpu_fast_omit(batch_size_, ssg_id, CacheFlag, super_batch_id, weightType);
    return nb_sample(batch_size_, ssg_id, CacheFlag, super_batch_id, weightType);
  }
};

#endif /* NTS_B200_SHADOW_NTSFASTSAMPLER_HPP */
