"""Host-side cold-row packer (csrc/pack_pool.h, used by the staging path of stage.cu): compiled alone with g++ and driven through
many jobs of every size -- rows land where they should, tiny jobs do not wake the pool, nothing hangs or double-counts."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include "pack_pool.h"
#include <stdio.h>
#include <stdlib.h>
int main() {
  const uint32_t V = 50000, F = 37, PITCH = 40;
  std::vector<float> table((size_t)V * PITCH);
  for (size_t i = 0; i < table.size(); i++) table[i] = (float)(i % 9973) * 0.5f;
  for (int n_threads : {1, 2, 7}) {
    PackPool pool;
    pool.start(n_threads);
    uint64_t x = 88172645463325252ull;
    for (int job = 0; job < 300; job++) {
      const uint32_t sizes[] = {0, 1, 127, 128, 129, 511, 512, 513, 4000, 30000};
      const uint32_t n = sizes[job % 10];
      std::vector<uint32_t> ids(n);
      for (auto &v : ids) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; v = (uint32_t)(x % V); }
      std::vector<float> dst((size_t)n * F + 1, -1.0f);
      pool.pack(table.data(), PITCH, ids.data(), dst.data(), n, F);
      for (uint32_t r = 0; r < n; r++)
        if (memcmp(&dst[(size_t)r * F], &table[(size_t)ids[r] * PITCH], F * sizeof(float))) { printf("MISMATCH job %d row %u\n", job, r); return 1; }
      if (dst[(size_t)n * F] != -1.0f) { printf("OVERRUN job %d\n", job); return 1; }
    }
    pool.shutdown();
  }
  printf("PACK_POOL_OK\n");
  return 0;
}
'''


def test_pack_pool_packs_every_job(tmp_path):
    src = tmp_path / "pp.cpp"
    src.write_text(SRC)
    exe = tmp_path / "pp"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "sample-based-gnn_b200", "csrc"), str(src), "-o", str(exe)])
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "PACK_POOL_OK" in r.stdout, r.stdout + r.stderr
