"""Host-side layout rule shared by the gather / aggregation launchers (csrc/common.cuh nb_pick_vec): which vector width a row layout
allows and how many columns a kernel may touch. Pure host code: compiled with nvcc and run on the CPU.

The contract it pins (include/nts_b200.h, "ROW PADDING"): dense rows are never over-run; padded rows move as whole 32-byte sectors
when the padding allows, else up to the vector width; mis-aligned bases or pitches fall back to narrower vectors."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include "common.cuh"
#include <stdio.h>
void nb_set_error(const char *, ...) {}
struct Case { uint32_t F; uintptr_t a; uint64_t pa; uintptr_t b; uint64_t pb; int with_feff; int vec; uint32_t feff; };
int main() {
  const uintptr_t A = 0x10000;   // 64 KB aligned
  const Case cases[] = {
    // F    base a  pitch a  base b  pitch b  f_eff?  -> vec, f_eff
    {602, A, 602, A, 602, 1, 2, 602},     // dense 602: 64-bit vectors, nothing beyond the row
    {602, A, 608, A, 608, 1, 4, 608},     // padded to 608: whole sectors
    {602, A, 604, A, 604, 1, 4, 604},     // padding only up to the vector width
    {602, A, 608, A, 604, 1, 4, 604},     // the tighter pitch decides
    {602, A, 608, A, 602, 1, 2, 602},     // one side dense -> 64-bit, dense length
    {128, A, 128, A, 128, 1, 4, 128},
    {100, A, 100, A, 100, 1, 4, 104 - 4}, // 100 = 25 float4: exact
    {100, A, 104, A, 104, 1, 4, 104},     // padded: 13 sectors
    {7,   A, 7,   A, 7,   1, 1, 7},
    {7,   A, 8,   A, 8,   1, 4, 8},
    {41,  A, 41,  A, 41,  1, 1, 41},
    {64,  A + 4, 64, A, 64, 1, 1, 64},    // base off by one float -> scalar
    {64,  A + 8, 64, A, 64, 1, 2, 64},    // base off by two floats -> 64-bit
    {602, A, 608, A, 608, 0, 2, 0},       // callers that cannot carry pad columns (push / epilogue): exact length only
    {608, A, 608, A, 608, 0, 4, 0},
  };
  int bad = 0;
  for (const Case &c : cases) {
    uint32_t fe = 0;
    const int v = nb_pick_vec(c.F, (const void *)c.a, c.pa, (const void *)c.b, c.pb, c.with_feff ? &fe : nullptr);
    const bool ok = v == c.vec && (!c.with_feff || fe == c.feff) && (!c.with_feff || (fe <= c.pa && fe <= c.pb && fe >= c.F && fe % v == 0));
    if (!ok) { printf("CASE F=%u pitch %llu/%llu: got vec %d f_eff %u, want %d %u\n", c.F, (unsigned long long)c.pa, (unsigned long long)c.pb, v, fe, c.vec, c.feff); bad++; }
  }
  printf(bad ? "PICK_VEC_FAILED\n" : "PICK_VEC_OK\n");
  return bad;
}
'''


def test_pick_vec_layout_rule(tmp_path):
    src = tmp_path / "pv.cu"
    src.write_text(SRC)
    exe = tmp_path / "pv"
    subprocess.check_call(["nvcc", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "sample-based-gnn_b200", "csrc"), str(src), "-o", str(exe)])
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "PICK_VEC_OK" in r.stdout, r.stdout + r.stderr
