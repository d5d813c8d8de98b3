// pack_pool.h -- host threads that pack scattered table rows into a contiguous (pinned) staging block. Plain C++ (no CUDA), so
// tests/test_pack_pool.py can exercise it without a GPU. Used by stage.cu.
#pragma once
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

// The staging block is read next by the GPU's copy engine, not by a core. Rows written with ordinary stores stay dirty in the
// cores' caches and the DMA read has to snoop every line out of them: measured on the B200 host (tools/h2d_probe.cu) a 46 MB
// pinned block copies at 55 GB/s when it was written with non-temporal stores and at 8.9 GB/s when 16 threads had just written
// it with ordinary stores. So rows are packed with streaming stores (SSE2, always present on x86-64).
#if defined(__x86_64__)
#include <emmintrin.h>
static inline void copy_row(float *dst, const float *src, uint32_t n) {
  uint32_t i = 0;
  for (; i < n && ((uintptr_t)(dst + i) & 15u); i++) _mm_stream_si32((int *)(dst + i), *(const int *)(src + i));
  for (; i + 4 <= n; i += 4) _mm_stream_ps(dst + i, _mm_loadu_ps(src + i));
  for (; i < n; i++) _mm_stream_si32((int *)(dst + i), *(const int *)(src + i));
}
static inline void store_fence() { _mm_sfence(); }
#else
static inline void copy_row(float *dst, const float *src, uint32_t n) { memcpy(dst, src, (size_t)n * sizeof(float)); }
static inline void store_fence() {}
#endif

struct PackPool {
  std::vector<std::thread> threads;
  std::mutex m;
  std::condition_variable cv_work, cv_done;
  uint64_t generation = 0;
  int running = 0;
  bool stop = false;
  // the job
  const float *table = nullptr;
  uint64_t table_pitch = 0;
  const uint32_t *ids = nullptr;
  float *dst = nullptr;
  uint32_t n = 0, F = 0;
  std::atomic<uint32_t> next{0};
  static constexpr uint32_t CHUNK = 128;  // rows per grab

  void work() {
    const size_t row_bytes = (size_t)F * sizeof(float);
    for (;;) {
      const uint32_t r0 = next.fetch_add(CHUNK);
      if (r0 >= n) return;
      const uint32_t r1 = r0 + CHUNK < n ? r0 + CHUNK : n;
      // random rows of a multi-GB table: every row is a chain of DRAM misses unless the lines of the rows ahead are requested early
      constexpr uint32_t AHEAD = 8;
      for (uint32_t r = r0; r < r1; r++) {
        if (r + AHEAD < r1) {
          const char *p = (const char *)(table + (uint64_t)ids[r + AHEAD] * table_pitch);
          for (size_t b = 0; b < row_bytes; b += 64) __builtin_prefetch(p + b, 0, 0);
        }
        copy_row(dst + (size_t)r * F, table + (uint64_t)ids[r] * table_pitch, F);
      }
      store_fence();
    }
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> g(m);
        cv_work.wait(g, [&] { return stop || generation != seen; });
        if (stop) return;
        seen = generation;
      }
      work();
      {
        std::lock_guard<std::mutex> g(m);
        if (--running == 0) cv_done.notify_all();
      }
    }
  }
  void start(int n_threads) {
    for (int i = 0; i + 1 < n_threads; i++) threads.emplace_back([this] { loop(); });  // the caller is the n-th worker
  }
  // packs dst[r,:] = table[ids[r],:] for r < n_rows with every thread of the pool plus the caller
  void pack(const float *table_, uint64_t pitch_, const uint32_t *ids_, float *dst_, uint32_t n_rows, uint32_t F_) {
    const bool wake = n_rows > 4 * CHUNK && !threads.empty();  // a tiny job is not worth waking anyone
    {
      std::lock_guard<std::mutex> g(m);
      table = table_; table_pitch = pitch_; ids = ids_; dst = dst_; n = n_rows; F = F_;
      next.store(0);
      running = wake ? (int)threads.size() : 0;
      if (wake) generation++;
    }
    if (wake) cv_work.notify_all();
    work();
    if (wake) {
      std::unique_lock<std::mutex> g(m);
      cv_done.wait(g, [&] { return running == 0; });  // every pool thread has seen this job and left it
    }
  }
  void shutdown() {
    { std::lock_guard<std::mutex> g(m); stop = true; }
    cv_work.notify_all();
    for (auto &t : threads) t.join();
    threads.clear();
  }
};

