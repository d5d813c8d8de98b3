// h2d_probe.cu -- does the host->device copy rate of a 46 MB pinned block depend on what the host cores are doing?
// (a) all cores idle between copies, (b) N threads spinning, (c) copy issued right after a multi-threaded memcpy burst.
// build: nvcc -O3 -o tools/h2d_probe tools/h2d_probe.cu -lpthread
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <chrono>
#include <thread>
#include <vector>
#include <immintrin.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
  const size_t bytes = 46u << 20;
  void *h, *d; CK(cudaHostAlloc(&h, bytes, cudaHostAllocDefault)); CK(cudaMalloc(&d, bytes)); memset(h, 1, bytes);
  cudaStream_t st; CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  auto copy_ms = [&](int reps, int sleep_us) {
    float tot = 0;
    for (int i = 0; i < reps; i++) {
      if (sleep_us) std::this_thread::sleep_for(std::chrono::microseconds(sleep_us));
      cudaEventRecord(a, st); cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st); cudaEventRecord(b, st); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b); tot += ms;
    }
    return tot / reps;
  };
  copy_ms(3, 0);
  printf("back-to-back copies            : %.3f ms (%.1f GB/s)\n", copy_ms(20, 0), bytes / copy_ms(20, 0) / 1e6);
  { float m = copy_ms(20, 3000); printf("3 ms idle before each copy     : %.3f ms (%.1f GB/s)\n", m, bytes / m / 1e6); }
  { float m = copy_ms(10, 30000); printf("30 ms idle before each copy    : %.3f ms (%.1f GB/s)\n", m, bytes / m / 1e6); }
  for (int n : {1, 4, 15}) {
    std::atomic<bool> stop{false};
    std::vector<std::thread> th;
    for (int i = 0; i < n; i++) th.emplace_back([&] { while (!stop.load(std::memory_order_relaxed)) { } });
    float m = copy_ms(10, 30000);
    printf("30 ms idle, %2d spinning threads : %.3f ms (%.1f GB/s)\n", n, m, bytes / m / 1e6);
    stop = true; for (auto &t : th) t.join();
  }
  // (d) the block is rewritten by 16 threads right before the copy: regular stores (lines stay dirty in the cores' caches and
  //     the DMA engine has to snoop them out) vs non-temporal stores (lines go to DRAM)
  for (int nt = 0; nt < 2; nt++) {
    float tot = 0, wtot = 0;
    for (int rep = 0; rep < 10; rep++) {
      std::vector<std::thread> th;
      const int T = 16;
      double w0 = now();
      for (int i = 0; i < T; i++) th.emplace_back([&, i] {
        float *p = (float *)h + (bytes / 4 / T) * i;
        const size_t n = bytes / 4 / T;
        if (nt) { for (size_t k = 0; k + 4 <= n; k += 4) _mm_stream_ps(p + k, _mm_set1_ps((float)(rep + k))); _mm_sfence(); }
        else for (size_t k = 0; k < n; k++) p[k] = (float)(rep + k);
      });
      for (auto &t : th) t.join();
      wtot += (float)((now() - w0) * 1e3);
      cudaEventRecord(a, st); cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st); cudaEventRecord(b, st); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b); tot += ms;
    }
    printf("rewritten by 16 threads with %s stores (%.2f ms) then copied: %.3f ms (%.1f GB/s)\n", nt ? "non-temporal" : "regular", wtot / 10, tot / 10, bytes / (tot / 10) / 1e6);
  }
  return 0;
}
