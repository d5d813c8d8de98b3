"""Source-level drop-in check: the reference's OWN toolkits/main.cpp (all 14 sample toolkits, hence every
core/*.hpp on the path) compiles unchanged against sample-based-gnn_b200/host/cuda/ntsCUDA.hpp -- the
header-only adaptor that re-implements cuda/ntsCUDA.hpp on top of the C ABI -- and the object then
depends on nb_* symbols instead of the reference's Cuda_Stream:: externals. Needs /root/reference
(build container only); skipped elsewhere."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("NTS_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources not present")
def test_reference_toolkits_compile_against_the_adaptor(tmp_path):
    import torch
    tdir = os.path.dirname(torch.__file__)
    obj = str(tmp_path / "main_adaptor.o")
    cmd = ["/usr/bin/g++", "-std=c++17", "-fopenmp", "-O0", "-march=x86-64-v3", "-w", "-DCUDA_ENABLE",
           f"-I{ROOT}/sample-based-gnn_b200/host", f"-I{ROOT}/sample-based-gnn_b200/host/cuda", f"-I{ROOT}/include",
           f"-I{ROOT}/oracle/shims", f"-I{REF}", f"-I{REF}/core", "-I/usr/local/cuda/include", f"-I{tdir}/include",
           f"-I{tdir}/include/torch/csrc/api/include", "-c", f"{REF}/toolkits/main.cpp", "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    syms = subprocess.run(["nm", "-C", obj], capture_output=True, text=True).stdout
    undefined = [l.split(" U ")[1] for l in syms.splitlines() if " U " in l]
    nb = {u for u in undefined if u.startswith("nb_")}
    assert len(nb) >= 25, nb                                      # the toolkits really call through the C ABI
    assert not [u for u in undefined if "Cuda_Stream::" in u]      # nothing left for the reference's CUDA library
    import __graft_entry__ as ge
    declared = set(ge.load_package()._capi.header_symbols())
    assert nb <= declared, nb - declared
