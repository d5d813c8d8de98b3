// cpp_e2e_bench.cpp -- the headline step driven from C++ through the adaptor header (sample-based-gnn_b200/host/cuda/ntsCUDA.hpp:
// the reference's own Cuda_Stream surface) and the C ABI, with HOST seeds in and the step's output back on the host every step.
// The same loop as bench.py's `e2e` arm without an interpreter between the calls: what a C++ toolkit pays per mini-batch.
//
//   cpp_e2e_bench <edge_pairs.bin> <V> <seeds.u32> <batch> <fanout a,b> <F0> <F1> <pitch> <steps> <warmup> <windows> [slots] [sampling streams] [bottom hop on its own stream: 0|1]
//
// edge_pairs.bin = raw (u32 src, u32 dst) pairs (the reference's EDGE_FILE format), seeds.u32 = training ids. Prints one JSON line.
// Step: sample (2 layers, pipeline slot i % slots, slots - 1 batches ahead, high-priority streams) -> bottom hop aggregated straight from the feature table
// (nb_aggregate_gathered_fwd_dyn) -> top hop forward (Cuda_Stream::Gather_By_Dst_From_Src_Spmm) -> top hop backward
// (Cuda_Stream::Gather_By_Src_From_Dst_Spmm) -> D2H of the [batch, F1] output into a pinned ring the host reads one step behind.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <unistd.h>
#include <vector>

#include "cuda/ntsCUDA.hpp"

#define CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e__)); exit(3); } } while (0)

static std::vector<uint32_t> read_u32(const char *path) {
  FILE *f = fopen(path, "rb");
  if (!f) { perror(path); exit(2); }
  fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
  std::vector<uint32_t> v(n / 4);
  if (fread(v.data(), 4, v.size(), f) != v.size()) exit(2);
  fclose(f);
  return v;
}

int main(int argc, char **argv) {
  if (argc < 12) { fprintf(stderr, "usage: see the header of tools/cpp_e2e_bench.cpp\n"); return 2; }
  const uint32_t V = (uint32_t)atol(argv[2]);
  std::vector<uint32_t> pairs = read_u32(argv[1]), seeds = read_u32(argv[3]);
  const uint32_t B = atoi(argv[4]);
  int fanout[2]; sscanf(argv[5], "%d,%d", &fanout[0], &fanout[1]);
  const uint32_t F0 = atoi(argv[6]), F1 = atoi(argv[7]), PITCH = atoi(argv[8]);
  const int K = atoi(argv[9]), W = atoi(argv[10]), R = atoi(argv[11]);
  constexpr int MAX_P = 8;
  const int P = argc > 12 ? std::max(2, std::min(MAX_P, atoi(argv[12]))) : 2;      // pipeline slots (the reference's PIPELINE_NUM)
  const int NS = argc > 13 ? std::max(1, std::min(P, atoi(argv[13]))) : 1;        // sampling streams: slot k samples on stream k % NS
  const bool AGG = argc > 14 ? atoi(argv[14]) != 0 : false;                       // the weight-free bottom hop on a stream of its own
  CK(cudaSetDevice(0));
  int lo_prio = 0, hi_prio = 0;
  CK(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
  cudaStream_t st_sample[MAX_P], st_train, st_agg;
  Cuda_Stream cs_sample[MAX_P], cs_train, cs_agg_own;    // the adaptor's class: same surface as the reference's
  for (int k = 0; k < NS; k++) {
    CK(cudaStreamCreateWithPriority(&st_sample[k], cudaStreamNonBlocking, hi_prio));
    cs_sample[k].setNewStream(st_sample[k]);
  }
  CK(cudaStreamCreateWithPriority(&st_train, cudaStreamNonBlocking, lo_prio));
  cs_train.setNewStream(st_train);
  st_agg = st_train;
  if (AGG) { CK(cudaStreamCreateWithPriority(&st_agg, cudaStreamNonBlocking, lo_prio)); cs_agg_own.setNewStream(st_agg); }
  Cuda_Stream &cs_agg = AGG ? cs_agg_own : cs_train;
  nb_graph *g = nullptr;
  NTS_B200_CHECK(nb_graph_create_from_pairs(cs_sample[0].ctx, V, pairs.size() / 2, pairs.data(), 0, &g));   // CSC built on the device
  std::vector<uint32_t>().swap(pairs);
  nb_sampler *smp[MAX_P];
  for (int k = 0; k < P; k++)
    NTS_B200_CHECK(nb_sampler_create(cs_sample[k % NS].ctx, g, 2, fanout, B, NB_SAMPLER_BUILD_CSR | NB_SAMPLER_NO_BOTTOM_CSR, 0, &smp[k]));
  // feature table in HBM (row pitch PITCH floats), synthetic values
  const size_t cap_s0 = std::min<size_t>((size_t)B * fanout[0], V);
  float *table = (float *)cudaMallocGPU((long)V * PITCH * 4), *y1[MAX_P];
  for (int k = 0; k < P; k++) y1[k] = (float *)cudaMallocGPU((long)cap_s0 * PITCH * 4);   // one Y1 per slot: batch i+1's bottom hop runs beside batch i's top hop
  float *h1 = (float *)cudaMallocGPU((long)cap_s0 * F1 * 4), *y0 = (float *)cudaMallocGPU((long)B * F1 * 4);
  float *dy0 = (float *)cudaMallocGPU((long)B * F1 * 4), *dh1 = (float *)cudaMallocGPU((long)cap_s0 * F1 * 4);
  {
    std::vector<float> chunk((size_t)4096 * PITCH);
    uint32_t x = 12345u;
    for (size_t r0 = 0; r0 < V; r0 += 4096) {
      const size_t n = std::min<size_t>(4096, V - r0);
      for (size_t i = 0; i < n * PITCH; i++) { x = x * 1664525u + 1013904223u; chunk[i] = (i % PITCH) < F0 ? (float)(x >> 8) * (1.0f / 8388608.0f) - 1.0f : 0.0f; }
      CK(cudaMemcpy(table + r0 * PITCH, chunk.data(), n * PITCH * 4, cudaMemcpyHostToDevice));
    }
    CK(cudaMemset(h1, 0, cap_s0 * F1 * 4)); CK(cudaMemset(dy0, 0, (size_t)B * F1 * 4));
  }
  float *y0_host[2];
  for (int k = 0; k < 2; k++) y0_host[k] = (float *)cudaMallocPinned((long)B * F1 * 4);
  cudaEvent_t sampled[MAX_P], consumed[MAX_P], aggregated[MAX_P], y0_done[2];
  for (int k = 0; k < P; k++) {
    CK(cudaEventCreateWithFlags(&sampled[k], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&aggregated[k], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&consumed[k], cudaEventDisableTiming));
    CK(cudaEventRecord(consumed[k], st_train));
  }
  for (int k = 0; k < 2; k++) CK(cudaEventCreateWithFlags(&y0_done[k], cudaEventDisableTiming));
  const long n_steps = W + (long)R * K;
  auto seed_of = [&](long i) { return &seeds[(size_t)((i * (long)B) % (long)(seeds.size() - B))]; };
  auto issue = [&](long i) {   // sample batch i into slot i % P, asynchronously, on the slot's high-priority stream
    const int k = (int)(i % P);
    CK(cudaStreamWaitEvent(st_sample[k % NS], consumed[k], 0));
    NTS_B200_CHECK(nb_sampler_sample(smp[k], seed_of(i), B, 0, 0x5EED0004ull, (uint64_t)i, NB_WEIGHT_SUM, nullptr, 0xffffffffu, nullptr, 0));
    CK(cudaEventRecord(sampled[k], st_sample[k % NS]));
  };
  double checksum = 0.0, edges = 0.0;
  auto step = [&](long i) {
    const int k = (int)(i % P), r = (int)(i % 2);
    nb_layer_view lv[2];
    NTS_B200_CHECK(nb_sampler_wait(smp[k], lv));                  // host learns this batch's sizes
    if (i + P - 1 < n_steps) issue(i + P - 1);                    // batch i-1's slot is free: sample ahead while this batch is aggregated
    CK(cudaStreamWaitEvent(st_agg, sampled[k], 0));
    const nb_layer_view &top = lv[0], &bot = lv[1];
    NTS_B200_CHECK(nb_aggregate_gathered_fwd_dyn(cs_agg.ctx, table, PITCH, bot.gather_index, y1[k], bot.edge_weight_forward, bot.column_offset,
                                                 nullptr, bot.n_dst, F0, PITCH));
    if (AGG) { CK(cudaEventRecord(aggregated[k], st_agg)); CK(cudaStreamWaitEvent(st_train, aggregated[k], 0)); }
    cs_train.Gather_By_Dst_From_Src_Spmm(h1, y0, (float *)top.edge_weight_forward, (VertexId_CUDA *)top.row_indices, (VertexId_CUDA *)top.column_offset,
                                         top.n_src, 0, 0, 0, 0, top.n_edges, top.n_dst, F1, true, false);
    cs_train.Gather_By_Src_From_Dst_Spmm(dy0, dh1, (float *)top.edge_weight_backward, (VertexId_CUDA *)top.row_offset, (VertexId_CUDA *)top.column_indices,
                                         top.n_dst, 0, 0, 0, 0, top.n_edges, top.n_src, F1, true, false);
    CK(cudaEventRecord(consumed[k], st_train));
    CK(cudaMemcpyAsync(y0_host[r], y0, (size_t)B * F1 * 4, cudaMemcpyDeviceToHost, st_train));
    CK(cudaEventRecord(y0_done[r], st_train));
    if (i > 0) { CK(cudaEventSynchronize(y0_done[1 - r])); checksum += y0_host[1 - r][0]; }
    edges += (double)top.n_edges + (double)bot.n_edges;
  };
  for (long i = 0; i < P - 1 && i < n_steps; i++) issue(i);
  for (long i = 0; i < W; i++) step(i);
  std::vector<double> win_ms, win_edges;
  cudaEvent_t t0, t1;
  CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1));
  for (int w = 0; w < R; w++) {
    CK(cudaDeviceSynchronize());
    const double e0 = edges;
    CK(cudaEventRecord(t0, st_train));
    for (long i = W + (long)w * K; i < W + (long)(w + 1) * K; i++) step(i);
    CK(cudaEventSynchronize(y0_done[(W + (long)(w + 1) * K - 1) % 2]));
    CK(cudaEventRecord(t1, st_train));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, t0, t1));
    win_ms.push_back(ms); win_edges.push_back(edges - e0);
  }
  std::vector<int> order(R);
  for (int w = 0; w < R; w++) order[w] = w;
  std::sort(order.begin(), order.end(), [&](int a, int b) { return win_ms[a] < win_ms[b]; });
  const int mid = order[R / 2];
  printf("{\"host\": \"C++ (adaptor header + C ABI)\", \"pipeline_slots\": %d, \"sampling_streams\": %d, \"bottom_hop_on_its_own_stream\": %s, \"value\": %.1f, \"unit\": \"edges/s\", \"ms_per_step\": %.6f, \"windows_ms_per_step\": [",
         P, NS, AGG ? "true" : "false", win_edges[mid] / (win_ms[mid] * 1e-3), win_ms[mid] / K);
  for (int w = 0; w < R; w++) printf("%s%.5f", w ? ", " : "", win_ms[w] / K);
  printf("], \"h2d_bytes_per_step\": %u, \"d2h_bytes_per_step\": %u, \"checksum\": %.6g}\n", B * 4 + 64, B * F1 * 4 + 3 * 32, checksum);
  fflush(stdout);
  _exit(0);
}
