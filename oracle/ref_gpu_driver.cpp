/* oracle/ref_gpu_driver.cpp -- drives the UNMODIFIED reference GPU path (TEST / MEASUREMENT INFRASTRUCTURE ONLY).
 *
 * This translation unit contains no algorithm of its own. It is compiled against the reference's own headers where they lie
 * under /root/reference and linked with the reference's own CUDA library source cuda/ntsCUDAGraphOP.cu (built for sm_100 by
 * oracle/Makefile, target _ref/ref_gpu_driver): the kernels it launches are the reference's (cuda/ntsCUDATransferKernel.cuh,
 * ntsCUDAFuseKernel.cuh, ntsCUDADistKernel.cuh) and the aggregation is the reference's cuSPARSE SpMM wrapper. It needs a GPU.
 *
 * Modes
 *   bench <edge_file> <V> <seed_file> <batch> <fanout a,b> <F0> <F1> <batches> <warmup> [gpu_sampler]
 *       the reference's mini-batch path of BASELINE config 2 (toolkits/GCN_SAMPLE_GPU.hpp:289-397: CPU sampling, GPU gather and
 *       aggregation) on the same graph / seeds as bench.py, stage by stage, in the toolkit's own call order:
 *         FastSampler::sample_fast(to_gpu)                core/ntsFastSampler.hpp:962-1140 (+ sampCSC::copy_data_to_device_async)
 *         Cuda_Stream::zero_copy_feature_move_gpu         cuda/ntsCUDAGraphOP.cu:1711-1729  (table in pinned host memory = the
 *                                                         reference layout, and the same kernel on an HBM-resident copy)
 *         Cuda_Stream::Gather_By_Dst_From_Src_Spmm        :425-587  (cuSPARSE SpMM, bottom hop F0 and top hop F1)
 *         Cuda_Stream::Gather_By_Src_From_Dst_Spmm        :901-1042 (cuSPARSE SpMM over the CSR, top hop backward:
 *                                                         SingleGPUSampleGraphOp::backward, core/ntsSingleGPUSampleGraphOp.hpp:126-176)
 *       each timed with CUDA events on the Cuda_Stream's stream (sampling: host wall clock around the call + stream
 *       synchronise). With a trailing `gpu_sampler` argument the batch is sampled by FastSampler::sample_gpu_fast instead
 *       (GCN_SAMPLE_ALLGPU's path; on the B200 box that kernel chain dies with an illegal memory access -- its |V|-sized mark
 *       array is cudaMalloc'd and never cleared, SURVEY.md section 8 quirks -- which is why it is not the default).
 *       Prints one JSON line.
 *   gat <in.bin> <out.bin>
 *       the reference's GAT edge kernels on a given sampled layer (record format of oracle/refio.py; arrays column_offset,
 *       row_indices, dst_local_id, h [S,F], att [2F], dout [V,F]; header word 1 = F, word 2 = S):
 *         Scatter_Src_Dst_to_Msg -> (libtorch: mm, leaky_relu 0.2) -> Edge_Softmax_Forward_Norm_Block -> (libtorch: mul) ->
 *         Gather_Msg_to_Dst, and the backward chain Scatter_Dst_to_Msg -> Edge_Softmax_Backward_Block -> Gather_Msg_To_Src_Dst
 *       exactly the op sequence of toolkits/GAT_SAMPLE_ALL_MULTI.hpp:383-464 / core/ntsPushdownGraphOp.hpp:490-747.
 *       Every intermediate is written to <out.bin>: the golden vectors of tests/golden/gat_*.npz.
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
#include <map>
#include <cuda_runtime.h>
#include "core/ntsMiniBatchGraphOp.hpp"

static std::vector<int> parse_fanout(const std::string &s) {
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
#define CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e__)); exit(3); } } while (0)

struct Timer {   /* CUDA events on the Cuda_Stream's stream */
  cudaEvent_t a, b; cudaStream_t st; double total = 0; int n = 0;
  explicit Timer(cudaStream_t s) : st(s) { CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); }
  void start() { CK(cudaEventRecord(a, st)); }
  void stop(bool count) { CK(cudaEventRecord(b, st)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (count) { total += ms; n++; } }
  double mean() const { return n ? total / n : 0; }
};

static int run_bench(int argc, char **argv) {
  if (argc < 11) return 2;
  const char *edge_file = argv[2]; VertexId V = atoi(argv[3]);
  std::vector<VertexId> seeds = read_u32(argv[4]);
  int batch = atoi(argv[5]); std::vector<int> fanout = parse_fanout(argv[6]);
  int F0 = atoi(argv[7]), F1 = atoi(argv[8]); int nb = atoi(argv[9]), warm = atoi(argv[10]);
  int L = fanout.size();
  CK(cudaSetDevice(0));
  Graph<Empty> *g = new Graph<Empty>();
  g->filename = edge_file; g->vertices = V;
  g->edges = file_size(edge_file) / (2 * sizeof(VertexId));
  g->partitions = 1; g->partition_id = 0; g->owned_vertices = V;
  g->partition_offset = new VertexId[2]; g->partition_offset[0] = 0; g->partition_offset[1] = V;
  g->gnnctx = new GNNContext(); g->gnnctx->layer_size = std::vector<int>(L + 1, F1); g->gnnctx->layer_size[0] = F0;
  g->config->up_degree = false;
  g->config->batch_size = batch;
  g->out_degree_for_backward = new VertexId[V](); g->in_degree_for_backward = new VertexId[V]();
  {
    std::vector<VertexId> ed = read_u32(edge_file);
    for (size_t i = 0; i + 1 < ed.size(); i += 2) { g->out_degree_for_backward[ed[i]]++; g->in_degree_for_backward[ed[i + 1]]++; }
    for (VertexId v = 0; v < V; v++) { if (g->in_degree_for_backward[v] < 1) g->in_degree_for_backward[v] = 1;
                                       if (g->out_degree_for_backward[v] < 1) g->out_degree_for_backward[v] = 1; }
  }
  FullyRepGraph *full = new FullyRepGraph(g);
  full->ReadRepGraphFromRawFile();     /* adjacency in mapped pinned host memory, core/FullyRepGraph.hpp:727 */
  const bool gpu_sampler = argc > 11 && std::string(argv[11]) == "gpu_sampler";
  Cuda_Stream *cs = new Cuda_Stream[1];
  FastSampler *sampler = gpu_sampler ? new FastSampler(full, seeds, L, batch, fanout, 1, cs)              /* GPU ctor, :125-176 */
                                     : new FastSampler(g, full, seeds, L, fanout, batch, true, 0, 1, cs); /* CPU sampler, to_gpu (GCN_SAMPLE_GPU.hpp:475) */
  cudaStream_t st = cs[0].stream;
  /* feature table: all ones (FEATURE_FILE:random, core/ntsDataloador.hpp:846-850) in pinned mapped host memory like GNNDatum's
   * (core/ntsDataloador.hpp:187), plus an HBM copy to time the same kernel without PCIe */
  float *host_table = (float *)cudaMallocPinned((long)V * F0 * sizeof(float));
  for (size_t i = 0; i < (size_t)V * F0; i++) host_table[i] = 1.0f;
  float *host_table_dev = (float *)getDevicePointer(host_table);
  float *hbm_table = (float *)cudaMallocGPU((long)V * F0 * sizeof(float));
  CK(cudaMemcpy(hbm_table, host_table, (size_t)V * F0 * sizeof(float), cudaMemcpyHostToDevice));
  size_t cap_s = (size_t)batch; for (int i = 0; i < L; i++) cap_s *= (size_t)fanout[i]; if (cap_s > V) cap_s = V;
  size_t cap_top = std::min<size_t>((size_t)batch * fanout[0], V);
  float *x0 = (float *)cudaMallocGPU((long)cap_s * F0 * sizeof(float));
  float *y1 = (float *)cudaMallocGPU((long)cap_top * F0 * sizeof(float));
  float *h1 = (float *)cudaMallocGPU((long)cap_top * F1 * sizeof(float));
  float *y0 = (float *)cudaMallocGPU((long)batch * F1 * sizeof(float));
  float *dy0 = (float *)cudaMallocGPU((long)batch * F1 * sizeof(float));
  float *dh1 = (float *)cudaMallocGPU((long)cap_top * F1 * sizeof(float));
  CK(cudaMemset(h1, 0, cap_top * F1 * sizeof(float))); CK(cudaMemset(dy0, 0, (size_t)batch * F1 * sizeof(float)));
  Timer t_gh(st), t_gd(st), t_f0(st), t_f1(st), t_b(st);
  double t_sample = 0; int done = 0; uint64_t edges = 0, E1 = 0, S1 = 0, V1 = 0;
  FILE *devnull = fopen("/dev/null", "w");
  for (int b = 0; b < nb + warm && sampler->work_offset < sampler->work_range[1]; b++) {
    const bool timed = b >= warm;
    fflush(stdout); int saved = dup(1); dup2(fileno(devnull), 1);   /* the sampler printf()s */
    CK(cudaStreamSynchronize(st));
    double s0 = get_time();
    SampledSubgraph *sg = gpu_sampler ? sampler->sample_gpu_fast(batch, 0) : sampler->sample_fast(batch, 0);
    CK(cudaStreamSynchronize(st));
    double s1 = get_time();
    fflush(stdout); dup2(saved, 1); close(saved);
    sampCSC *bot = sg->sampled_sgs[L - 1], *top = sg->sampled_sgs[0];
    if (bot->src_size > cap_s || top->src_size > cap_top) { fprintf(stderr, "capacity\n"); return 3; }
    t_gh.start(); cs[0].zero_copy_feature_move_gpu(x0, host_table_dev, bot->dev_source, F0, bot->src_size); t_gh.stop(timed);
    t_gd.start(); cs[0].zero_copy_feature_move_gpu(x0, hbm_table, bot->dev_source, F0, bot->src_size); t_gd.stop(timed);
    t_f0.start();
    float *bw = gpu_sampler ? bot->dev_e_w() : bot->dev_e_w_f(), *tw = gpu_sampler ? top->dev_e_w() : top->dev_e_w_f();
    cs[0].Gather_By_Dst_From_Src_Spmm(x0, y1, bw, bot->dev_r_i(), bot->dev_c_o(), bot->src_size, 0, 0, 0, 0, bot->e_size, bot->v_size, F0, true, false);
    t_f0.stop(timed);
    t_f1.start();
    cs[0].Gather_By_Dst_From_Src_Spmm(h1, y0, tw, top->dev_r_i(), top->dev_c_o(), top->src_size, 0, 0, 0, 0, top->e_size, top->v_size, F1, true, false);
    t_f1.stop(timed);
    t_b.start();
    if (gpu_sampler) cs[0].Push_From_Dst_To_Src_Spmm(dy0, dh1, tw, top->dev_r_i(), top->dev_c_o(), top->src_size, 0, 0, 0, 0, top->e_size, top->v_size, F1, true, false);
    else cs[0].Gather_By_Src_From_Dst_Spmm(dy0, dh1, top->dev_e_w_b(), top->dev_r_o(), top->dev_c_i(), top->v_size, 0, 0, 0, 0, top->e_size, top->src_size, F1, true, false);
    t_b.stop(timed);
    if (timed) {
      t_sample += s1 - s0; done++;
      for (int i = 0; i < L; i++) edges += sg->sampled_sgs[i]->e_size;
      E1 += bot->e_size; S1 += bot->src_size; V1 += bot->v_size;
    }
  }
  printf("{\"sampler\": \"%s\", \"batches\": %d, \"sample_ms\": %.4f, \"gather_host_table_ms\": %.4f, \"gather_hbm_table_ms\": %.4f, \"spmm_fwd_F0_ms\": %.4f, "
         "\"spmm_fwd_F1_ms\": %.4f, \"spmm_bwd_F1_ms\": %.4f, \"edges\": %lu, \"E1\": %lu, \"S1\": %lu, \"V1\": %lu}\n",
         gpu_sampler ? "sample_gpu_fast" : "sample_fast (CPU, OpenMP) + copy_data_to_device_async", done, done ? t_sample / done * 1e3 : 0.0, t_gh.mean(), t_gd.mean(), t_f0.mean(), t_f1.mean(), t_b.mean(),
         (unsigned long)edges, (unsigned long)E1, (unsigned long)S1, (unsigned long)V1);
  fflush(stdout);
  _exit(0);   /* the reference's destructors double-free device arenas (core/FullyRepGraph.hpp:182-183) */
}

/* ---- record files (oracle/refio.py) ---- */
struct Rec { std::map<std::string, std::vector<uint32_t>> u; std::map<std::string, std::vector<float>> f; uint32_t hdr[4]; };
static Rec read_rec(const char *path) {
  Rec r; FILE *fp = fopen(path, "rb"); if (!fp) { perror(path); exit(2); }
  if (fread(r.hdr, 4, 4, fp) != 4) exit(2);
  while (true) {
    char nm[17] = {0}; uint32_t dt; uint64_t n;
    if (fread(nm, 1, 16, fp) != 16) break;
    if (fread(&dt, 4, 1, fp) != 1 || fread(&n, 8, 1, fp) != 1) break;
    if (!strcmp(nm, "end")) break;
    if (dt == 0) { std::vector<uint32_t> v(n); if (n && fread(v.data(), 4, n, fp) != n) exit(2); r.u[nm] = v; }
    else { std::vector<float> v(n); if (n && fread(v.data(), 4, n, fp) != n) exit(2); r.f[nm] = v; }
  }
  fclose(fp); return r;
}
struct Out {
  FILE *f;
  void tag(const char *name, uint32_t dtype, uint64_t n) { char nm[16] = {0}; strncpy(nm, name, 15); fwrite(nm, 1, 16, f); fwrite(&dtype, 4, 1, f); fwrite(&n, 8, 1, f); }
  void f32(const char *name, const torch::Tensor &t) { torch::Tensor c = t.contiguous().cpu(); tag(name, 1, c.numel()); fwrite(c.data_ptr<float>(), 4, c.numel(), f); }
};

static int run_gat(int argc, char **argv) {
  if (argc < 4) return 2;
  Rec in = read_rec(argv[2]);
  const uint32_t F = in.hdr[1], S = in.hdr[2];
  std::vector<uint32_t> &co = in.u["column_offset"], &ri = in.u["row_indices"], &dl = in.u["dst_local_id"];
  const uint32_t Vd = co.size() - 1, E = ri.size();
  CK(cudaSetDevice(0));
  Cuda_Stream *cs = new Cuda_Stream();
  auto dev_u32 = [&](std::vector<uint32_t> &v) { VertexId_CUDA *p; allocate_gpu_edge(&p, v.size() ? v.size() : 1); move_bytes_in(p, v.data(), v.size() * 4); return p; };
  VertexId_CUDA *d_co = dev_u32(co), *d_ri = dev_u32(ri), *d_dl = dev_u32(dl);
  auto opt = torch::TensorOptions().dtype(torch::kFloat32).device(torch::kCUDA, 0);
  auto host = [&](std::vector<float> &v, std::vector<int64_t> shape) { return torch::from_blob(v.data(), shape, torch::kFloat32).clone().to(torch::kCUDA); };
  torch::Tensor h = host(in.f["h"], {(long)S, (long)F}), att = host(in.f["att"], {2 * (long)F, 1}), dout = host(in.f["dout"], {(long)Vd, (long)F});
  cudaStream_t st = cs->stream;
  auto sync = [&]() { CK(cudaStreamSynchronize(st)); CK(cudaDeviceSynchronize()); };
  /* forward: BatchGPUSrcDstScatterOp -> edge NN -> BatchGPUEdgeSoftMax -> mul -> BatchGPUAggregateDst */
  torch::Tensor e_msg = torch::zeros({(long)E, 2 * (long)F}, opt);
  sync();
  cs->Scatter_Src_Dst_to_Msg(e_msg.data_ptr<float>(), h.data_ptr<float>(), d_ri, d_co, Vd, F, d_dl);
  sync();
  torch::Tensor s = e_msg.mm(att);                       /* [E,1], Parameter::forward = x.mm(W) */
  torch::Tensor m = torch::leaky_relu(s, 0.2);
  torch::Tensor a = torch::zeros({(long)E, 1}, opt), cached = torch::zeros({(long)E, 1}, opt);
  sync();
  cs->Edge_Softmax_Forward_Norm_Block(a.data_ptr<float>(), m.data_ptr<float>(), cached.data_ptr<float>(), d_ri, d_co, Vd, 1);
  sync();
  torch::Tensor e_msg_out = (e_msg.slice(1, 0, F, 1) * a).contiguous();
  torch::Tensor nbr = torch::zeros({(long)Vd, (long)F}, opt);
  sync();
  cs->Gather_Msg_to_Dst(nbr.data_ptr<float>(), e_msg_out.data_ptr<float>(), d_ri, d_co, Vd, F);
  sync();
  /* backward of the same chain */
  torch::Tensor d_e_msg_out = torch::zeros({(long)E, (long)F}, opt);
  cs->Scatter_Dst_to_Msg(d_e_msg_out.data_ptr<float>(), dout.data_ptr<float>(), d_ri, d_co, Vd, F);
  sync();
  torch::Tensor d_a = (d_e_msg_out * e_msg.slice(1, 0, F, 1)).sum(1, true).contiguous();
  torch::Tensor d_m = torch::zeros({(long)E, 1}, opt);
  sync();
  cs->Edge_Softmax_Backward_Block(d_m.data_ptr<float>(), d_a.data_ptr<float>(), cached.data_ptr<float>(), d_ri, d_co, Vd, 1);
  sync();
  torch::Tensor d_s = d_m * torch::where(s > 0, torch::ones_like(s), torch::full_like(s, 0.2));
  torch::Tensor d_att = e_msg.t().mm(d_s);               /* [2F,1] */
  torch::Tensor d_e_msg = d_s.mm(att.t()).contiguous();  /* [E,2F] */
  d_e_msg.slice(1, 0, F, 1) += d_e_msg_out * a;
  torch::Tensor dh = torch::zeros({(long)S, (long)F}, opt);
  sync();
  cs->Gather_Msg_To_Src_Dst(dh.data_ptr<float>(), d_e_msg.data_ptr<float>(), d_ri, d_co, Vd, F, d_dl);
  sync();
  Out o{fopen(argv[3], "wb")};
  uint32_t hdr[4] = {0x4e545352u, F, S, E};
  fwrite(hdr, 4, 4, o.f);
  o.f32("e_msg", e_msg); o.f32("score_pre", s); o.f32("m", m); o.f32("alpha", a); o.f32("cached", cached);
  o.f32("e_msg_out", e_msg_out); o.f32("out", nbr);
  o.f32("d_e_msg_out", d_e_msg_out); o.f32("d_a", d_a); o.f32("d_m", d_m); o.f32("d_e_msg", d_e_msg);
  o.f32("dh", dh); o.f32("datt", d_att);
  o.tag("end", 0, 0);
  fclose(o.f);
  fflush(stdout);
  _exit(0);
}

int main(int argc, char **argv) {
  if (argc < 2) { fprintf(stderr, "usage: see header\n"); return 2; }
  std::string mode = argv[1];
  if (mode == "bench") return run_bench(argc, argv);
  if (mode == "gat") return run_gat(argc, argv);
  return 2;
}
