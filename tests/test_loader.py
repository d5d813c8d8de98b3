"""CPU test of the feature / label / mask loader (csrc/loader.cu, host code of libnts_b200.so; no GPU involved): its values must be
bit-identical to what the reference's reader produces -- `ifstream >> float` token by token (core/ntsDataloador.hpp:999-1063) --
which a few lines of C++ compiled here with g++ reproduce as the checker."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import __graft_entry__ as ge

CHECKER = r"""
#include <fstream>
#include <cstdio>
#include <cstdint>
#include <string>
#include <vector>
int main(int argc, char **argv) {            // argv: feature_file V F out.bin   (the reference's loop shape: id, then F floats)
  std::ifstream in(argv[1]); unsigned V = atoi(argv[2]), F = atoi(argv[3]);
  std::vector<float> t((size_t)V * F, 0.f); unsigned id;
  while (in >> id) for (unsigned i = 0; i < F; i++) in >> t[(size_t)id * F + i];
  FILE *f = fopen(argv[4], "wb"); fwrite(t.data(), 4, t.size(), f); fclose(f); return 0;
}
"""


@pytest.fixture(scope="module")
def pkg():
    if not os.path.exists(os.path.join(ge.PKG_DIR, "lib", "libnts_b200.so")):
        ge.build()
    return ge.load_package()


def test_text_loader_matches_iostream_and_binary_cache(pkg, tmp_path):
    lib, check = pkg._capi.lib(), pkg._capi.check
    V, F = 300, 37
    rng = np.random.default_rng(5)
    vals = rng.standard_normal((V, F)) * 10.0 ** rng.integers(-8, 8, (V, F))
    order = rng.permutation(V)
    fmt = ["%.9g", "%.17g", "%e", "%f", "%.3f"]
    with open(tmp_path / "feat.txt", "w") as f:
        for v in order:
            f.write(str(v) + " " + " ".join(fmt[(v + k) % len(fmt)] % vals[v, k] for k in range(F)) + ("\r\n" if v % 7 == 0 else "\n"))
        f.write("\n")
    words = ["train", "eval", "val", "test", "unknown"]
    labels = rng.integers(0, 47, V)
    with open(tmp_path / "label.txt", "w") as f:
        f.writelines(f"{v} {labels[v]}\n" for v in order)
    with open(tmp_path / "mask.txt", "w") as f:
        f.writelines(f"{v} {words[v % 5]}\n" for v in order)
    (tmp_path / "chk.cpp").write_text(CHECKER)
    subprocess.check_call(["g++", "-O1", "-o", str(tmp_path / "chk"), str(tmp_path / "chk.cpp")])
    subprocess.check_call([str(tmp_path / "chk"), str(tmp_path / "feat.txt"), str(V), str(F), str(tmp_path / "ref.bin")])
    ref = np.fromfile(tmp_path / "ref.bin", np.float32).reshape(V, F)
    out = np.zeros((V, F), np.float32)
    hit = C.c_int(-1)
    path = str(tmp_path / "feat.txt").encode()
    check(lib.nb_read_feature_table(path, V, F, 0, V, out.ctypes.data, 1, C.byref(hit)))
    assert hit.value == 0 and np.array_equal(out.view(np.uint32), ref.view(np.uint32))      # parsed text == operator>>, bit for bit
    assert os.path.exists(str(tmp_path / "feat.txt") + ".nb_f32")
    out2 = np.zeros((V, F), np.float32)
    check(lib.nb_read_feature_table(path, V, F, 0, V, out2.ctypes.data, 1, C.byref(hit)))
    assert hit.value == 1 and np.array_equal(out2.view(np.uint32), ref.view(np.uint32))     # second call: the binary cache
    part = np.zeros((100, F), np.float32)                                                    # a partition's id range, from the cache
    check(lib.nb_read_feature_table(path, V, F, 50, 150, part.ctypes.data, 1, C.byref(hit)))
    assert hit.value == 1 and np.array_equal(part, ref[50:150])
    part2 = np.zeros((100, F), np.float32)                                                   # ... and from the text
    check(lib.nb_read_feature_table(path, V, F, 50, 150, part2.ctypes.data, 0, C.byref(hit)))
    assert hit.value == 0 and np.array_equal(part2, ref[50:150])
    lab = np.full(V, -1, np.int64)
    msk = np.full(V, -1, np.int32)
    check(lib.nb_read_label_mask(str(tmp_path / "label.txt").encode(), str(tmp_path / "mask.txt").encode(), 0, V, lab.ctypes.data, msk.ctypes.data))
    assert np.array_equal(lab, labels)
    assert np.array_equal(msk, np.array([{0: 0, 1: 1, 2: 1, 3: 2, 4: 3}[v % 5] for v in range(V)], np.int32))
    (tmp_path / "bad.txt").write_text("0 1.0 2.0\n1 3.0 oops\n")
    bad = np.zeros((2, 2), np.float32)
    assert lib.nb_read_feature_table(str(tmp_path / "bad.txt").encode(), 2, 2, 0, 2, bad.ctypes.data, 0, None) != 0
    assert lib.nb_read_feature_table(b"/nonexistent/file", 2, 2, 0, 2, bad.ctypes.data, 0, None) != 0
