/* oracle/ref_driver.cpp -- drives the UNMODIFIED reference CPU path (TEST INFRASTRUCTURE ONLY).
 *
 * This translation unit contains no algorithm of its own. It #includes the
 * reference headers where they lie under /root/reference and calls, in the
 * reference's own order (toolkits/GCN_CPU_SAMPLE.hpp:192-235):
 *
 *   FullyRepGraph::ReadRepGraphFromRawFile      core/FullyRepGraph.hpp:724-798
 *   FastSampler::sample_fast                    core/ntsFastSampler.hpp:962-1140
 *     (-> sampCSC::csc_to_csr  core/coocsc.hpp:82-111,
 *         sampCSC::WeightCompute core/coocsc.hpp:301-324,
 *         nts_norm_degree core/ntsBaseOp.hpp:652-657)
 *   nts::op::get_feature                        core/ntsMiniBatchGraphOp.hpp:45-60
 *   MiniBatchFuseOp::forward / backward         core/ntsMiniBatchGraphOp.hpp:143-270
 *
 * Build recipe: oracle/Makefile (target _ref/ref_driver). MPI / libnuma / boost
 * are absent from the image; oracle/shims/ supplies single-rank stand-ins and
 * oracle/ref_stubs.cpp supplies malloc-backed stand-ins for the few CUDA-side
 * allocation helpers the CPU path touches (cudaMallocPinned, Cuda_Stream ctor).
 *
 * Modes
 *   record <edge_file> <V> <seed_file> <batch> <fanout a,b> <F> <out.bin> [up_degree] [weight: sum|mean|none]
 *       samples every batch of <seed_file> (u32 ids) and appends, per batch and per
 *       layer, every sampCSC array plus X0 / forward / backward tensors to <out.bin>
 *       in the record format documented in oracle/refio.py.
 *   bench  <edge_file> <V> <seed_file> <batch> <fanout> <F0> <F1> <batches> <warmup>
 *       times sample_fast / get_feature / forward / backward over the first
 *       <batches> batches after <warmup>; prints one JSON line.
 *
 * The feature table is synthetic and formula-defined (feat(v,j), below) so that
 * tests can regenerate it without shipping a 15 MB file.
 */
#include <random>
#include <chrono>
#include <algorithm>
#include <unordered_map>
#include <fstream>
#include <execution>
#include <cstdio>
#include <cstdint>
#include <string>
#include <vector>
#include "core/ntsMiniBatchGraphOp.hpp"

static float feat(uint32_t v, uint32_t j) {
  /* exactly representable in fp32: a small signed multiple of 1/64 */
  uint32_t h = v * 2654435761u + j * 40503u + 12345u;
  h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
  return (float)((int)(h % 257u) - 128) / 64.0f;
}

static std::vector<int> parse_fanout(const std::string &s) { /* comma separated, e.g. 25,10 or -1,10 */
  std::vector<int> f; size_t p = 0;
  while (p <= s.size()) { size_t q = s.find(',', p); if (q == std::string::npos) q = s.size();
    f.push_back(atoi(s.substr(p, q - p).c_str())); p = q + 1; }
  return f;
}

static std::vector<VertexId> read_u32(const char *path) {
  FILE *f = fopen(path, "rb"); if (!f) { perror(path); exit(2); }
  fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
  std::vector<VertexId> v(n / 4); if (fread(v.data(), 4, v.size(), f) != v.size()) exit(2); fclose(f); return v;
}

struct Out {
  FILE *f;
  void tag(const char *name, uint32_t dtype, uint64_t n) { /* dtype 0=u32 1=f32 */
    char nm[16] = {0}; strncpy(nm, name, 15); fwrite(nm, 1, 16, f); fwrite(&dtype, 4, 1, f); fwrite(&n, 8, 1, f); }
  void u32(const char *name, const VertexId *p, uint64_t n) { tag(name, 0, n); fwrite(p, 4, n, f); }
  void f32(const char *name, const float *p, uint64_t n) { tag(name, 1, n); fwrite(p, 4, n, f); }
};

struct Env {
  Graph<Empty> *graph; FullyRepGraph *full; torch::Tensor feature;
};

static Env setup(const char *edge_file, VertexId V, const std::vector<int> &layer_size, bool up_degree, bool need_feature) {
  Env e;
  Graph<Empty> *g = new Graph<Empty>();
  g->filename = edge_file;
  g->vertices = V;
  long bytes = file_size(edge_file);
  g->edges = bytes / (2 * sizeof(VertexId));
  g->partitions = 1; g->partition_id = 0; g->owned_vertices = V;
  g->partition_offset = new VertexId[2]; g->partition_offset[0] = 0; g->partition_offset[1] = V;
  g->gnnctx = new GNNContext(); g->gnnctx->layer_size = layer_size;
  g->config->up_degree = up_degree;
  /* degrees exactly as Graph::load_directed + generate_backward_structure leave them
   * (core/graph.hpp:1305-1424, clamp at :4525-4530): per-edge counts, clamped to >= 1 */
  g->out_degree_for_backward = new VertexId[V](); g->in_degree_for_backward = new VertexId[V]();
  {
    std::vector<VertexId> ed = read_u32(edge_file);
    for (size_t i = 0; i + 1 < ed.size(); i += 2) { g->out_degree_for_backward[ed[i]]++; g->in_degree_for_backward[ed[i + 1]]++; }
    for (VertexId v = 0; v < V; v++) { if (g->in_degree_for_backward[v] < 1) g->in_degree_for_backward[v] = 1;
                                       if (g->out_degree_for_backward[v] < 1) g->out_degree_for_backward[v] = 1; }
  }
  e.graph = g;
  e.full = new FullyRepGraph(g);
  e.full->ReadRepGraphFromRawFile();
  if (need_feature) {
    int F = layer_size[0];
    e.feature = torch::empty({(long)V, (long)F}, torch::kFloat32);
    float *p = e.feature.data_ptr<float>();
    for (VertexId v = 0; v < V; v++) for (int j = 0; j < F; j++) p[(size_t)v * F + j] = feat(v, j);
  }
  return e;
}

int main(int argc, char **argv) {
  if (argc < 2) { fprintf(stderr, "usage: see header\n"); return 2; }
  std::string mode = argv[1];
  if (mode == "record") {
    if (argc < 9) return 2;
    const char *edge_file = argv[2]; VertexId V = atoi(argv[3]);
    std::vector<VertexId> seeds = read_u32(argv[4]);
    int batch = atoi(argv[5]); std::vector<int> fanout = parse_fanout(argv[6]); int F = atoi(argv[7]);
    const char *out_path = argv[8];
    bool up_degree = argc > 9 && atoi(argv[9]) != 0;
    std::string wt = argc > 10 ? argv[10] : "sum";
    WeightType weightType = wt == "mean" ? WeightType::Mean : (wt == "none" ? WeightType::None : WeightType::Sum);
    int L = fanout.size();
    std::vector<int> layer_size(L + 1, F);
    Env e = setup(edge_file, V, layer_size, up_degree, true);
    FastSampler *sampler = new FastSampler(e.graph, e.full, seeds, L, fanout, batch, false);
    Out o{fopen(out_path, "wb")};
    uint32_t hdr[4] = {0x4e545352u /*"NTSR"*/, (uint32_t)L, (uint32_t)F, (uint32_t)up_degree};
    fwrite(hdr, 4, 4, o.f);
    o.u32("g_col_off", e.full->column_offset, (uint64_t)V + 1);
    o.u32("g_row_idx", e.full->row_indices, e.full->global_edges);
    o.u32("g_in_deg", e.graph->in_degree_for_backward, V);
    o.u32("g_out_deg", e.graph->out_degree_for_backward, V);
    while (sampler->work_offset < sampler->work_range[1]) {
      SampledSubgraph *sg = sampler->sample_fast(batch, weightType);
      o.tag("batch", 0, 0);
      for (int i = 0; i < L; i++) {
        sampCSC *c = sg->sampled_sgs[i];
        o.tag("layer", 0, 0);
        o.u32("destination", c->destination.data(), c->destination.size());
        o.u32("column_offset", c->column_offset.data(), c->column_offset.size());
        o.u32("sample_ans", c->sample_ans.data(), c->sample_ans.size());
        o.u32("source", c->source.data(), c->source.size());
        o.u32("row_indices", c->row_indices.data(), c->row_indices.size());
        o.u32("row_offset", c->row_offset.data(), c->row_offset.size());
        o.u32("column_indices", c->column_indices.data(), c->column_indices.size());
        o.f32("e_w_f", c->edge_weight_forward.data(), c->edge_weight_forward.size());
        o.f32("e_w_b", c->edge_weight_backward.data(), c->edge_weight_backward.size());
      }
      if (up_degree) { /* UP_DEGREE overwrites the graph's degree arrays per layer (core/FullyRepGraph.hpp:189-207);
                        * what is left after sample_fast() are the LAST layer's, which the CPU op then reads */
        o.u32("deg_in", e.graph->in_degree_for_backward, V);
        o.u32("deg_out", e.graph->out_degree_for_backward, V);
      }
      /* gather + aggregate with the degrees as the last layer left them, as the toolkit does */
      NtsVar X0 = nts::op::get_feature(sg->sampled_sgs[L - 1]->src(), e.feature, e.graph);
      o.f32("X0", X0.data_ptr<float>(), X0.numel());
      NtsVar X = X0;
      std::vector<NtsVar> ys;
      for (int l = 0; l < L; l++) {
        int hop = (L - 1) - l;
        nts::op::MiniBatchFuseOp op(sg, e.graph, hop);
        NtsVar Y = op.forward(X);
        char nm[16]; snprintf(nm, 16, "Y%d", hop); o.f32(nm, Y.data_ptr<float>(), Y.numel());
        /* deterministic dY = Y scaled, then backward through the same op */
        NtsVar dY = (Y * 0.5f + 0.25f).contiguous();
        NtsVar dX = op.backward(dY);
        snprintf(nm, 16, "dX%d", hop); o.f32(nm, dX.data_ptr<float>(), dX.numel());
        X = Y;
      }
    }
    o.tag("end", 0, 0);
    fclose(o.f);
    return 0;
  }
  if (mode == "bench") {
    if (argc < 11) return 2;
    const char *edge_file = argv[2]; VertexId V = atoi(argv[3]);
    std::vector<VertexId> seeds = read_u32(argv[4]);
    int batch = atoi(argv[5]); std::vector<int> fanout = parse_fanout(argv[6]);
    int F0 = atoi(argv[7]), F1 = atoi(argv[8]); int nb = atoi(argv[9]), warm = atoi(argv[10]);
    int L = fanout.size();
    std::vector<int> layer_size(L + 1, F1); layer_size[0] = F0;
    double t_setup = -get_time();
    Env e = setup(edge_file, V, layer_size, false, false);
    /* all-ones table = the reference's FEATURE_FILE:random (core/ntsDataloador.hpp:846-850) */
    e.feature = torch::ones({(long)V, (long)F0}, torch::kFloat32);
    t_setup += get_time();
    FastSampler *sampler = new FastSampler(e.graph, e.full, seeds, L, fanout, batch, false);
    double ts = 0, tg = 0, tf = 0, tb = 0; uint64_t edges = 0, rows = 0; int done = 0;
    std::vector<uint64_t> Es(L, 0), Ss(L, 0);
    FILE *devnull = fopen("/dev/null", "w");
    for (int b = 0; b < nb + warm && sampler->work_offset < sampler->work_range[1]; b++) {
      bool timed = b >= warm;
      fflush(stdout); int saved = dup(1); dup2(fileno(devnull), 1); /* sample_fast printf()s per call (:972) */
      double t0 = get_time();
      SampledSubgraph *sg = sampler->sample_fast(batch);
      double t1 = get_time();
      fflush(stdout); dup2(saved, 1); close(saved);
      NtsVar X0 = nts::op::get_feature(sg->sampled_sgs[L - 1]->src(), e.feature, e.graph);
      double t2 = get_time();
      /* bottom hop aggregates F0-wide rows, upper hops F1-wide (the dense layer between is libtorch, out of path) */
      std::vector<NtsVar> Y(L), dX(L);
      double fwd = 0, bwd = 0;
      NtsVar X = X0;
      for (int l = 0; l < L; l++) {
        int hop = (L - 1) - l;
        if (l > 0) X = torch::ones({(long)sg->sampled_sgs[hop]->src().size(), (long)F1}, torch::kFloat32);
        nts::op::MiniBatchFuseOp op(sg, e.graph, hop);
        double a = get_time(); Y[l] = op.forward(X); double c = get_time();
        NtsVar dY = torch::ones_like(Y[l]);
        /* NtsContext::self_backward stops before the first op on the tape (core/ntsContext.hpp:443:
         * `while (count > 1 || ...)`), so the bottom hop's backward into X0 never runs in the toolkit */
        double d = get_time(); if (l > 0) dX[l] = op.backward(dY); double f = get_time();
        fwd += c - a; bwd += f - d;
      }
      if (timed) {
        ts += t1 - t0; tg += t2 - t1; tf += fwd; tb += bwd; done++;
        for (int i = 0; i < L; i++) { edges += sg->sampled_sgs[i]->e_size; Es[i] += sg->sampled_sgs[i]->e_size; Ss[i] += sg->sampled_sgs[i]->src_size; }
        rows += sg->sampled_sgs[L - 1]->src_size;
      }
    }
    printf("{\"batches\": %d, \"threads\": %d, \"setup_s\": %.3f, \"sample_s\": %.6f, \"gather_s\": %.6f, \"fwd_s\": %.6f, \"bwd_s\": %.6f, \"edges\": %lu, \"rows\": %lu",
           done, sampler->ssg->threads, t_setup, ts, tg, tf, tb, (unsigned long)edges, (unsigned long)rows);
    for (int i = 0; i < L; i++) printf(", \"E%d\": %lu, \"S%d\": %lu", i, (unsigned long)Es[i], i, (unsigned long)Ss[i]);
    printf("}\n");
    return 0;
  }
  if (mode == "hotness") {
    /* hotness <edge_file> <V> <seed_file> <batch> <pipeline> <layers> <out.bin>
     * nts::op::preSample (core/ntsBaseOp.hpp:415-541), the overload the *_CACHE toolkits call: per super-batch
     * get_most_neighbor (:333-399) and the counts||ids file. The file name is derived from the edge file by the
     * reference itself (:420-428, which also forces cache_rate = 0.8 when the file does not exist yet). */
    if (argc < 9) return 2;
    const char *edge_file = argv[2]; VertexId V = atoi(argv[3]);
    std::vector<VertexId> seeds = read_u32(argv[4]);
    int batch = atoi(argv[5]), pipeline = atoi(argv[6]), layers = atoi(argv[7]);
    std::vector<int> layer_size(layers + 1, 4);
    Env e = setup(edge_file, V, layer_size, false, false);
    e.graph->config->edge_file = edge_file;
    e.graph->config->pre_sample_file = "";
    e.graph->config->batch_size = batch;
    e.graph->config->fanout_string = "x";
    std::vector<VertexId> batch_cache_num;
    VertexId top_cache_num = 0;
    std::vector<VertexId> ids = nts::op::preSample(seeds, batch, batch_cache_num, 0.5f, top_cache_num, layers, e.full, 1.0f, e.graph, pipeline);
    Out o{fopen(argv[8], "wb")};
    uint32_t hdr[4] = {0x4e545352u, (uint32_t)layers, (uint32_t)batch, (uint32_t)pipeline};
    fwrite(hdr, 4, 4, o.f);
    o.u32("counts", batch_cache_num.data(), batch_cache_num.size());
    o.u32("ids", ids.data(), ids.size());
    o.u32("top", &top_cache_num, 1);
    o.tag("end", 0, 0);
    fclose(o.f);
    return 0;
  }
  return 2;
}
