"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports
every symbol include/nts_b200.h declares, and fails loudly (no fallback) without a CUDA device."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import __graft_entry__ as ge


@pytest.fixture(scope="module")
def pkg():
    if not os.path.exists(os.path.join(ge.PKG_DIR, "lib", "libnts_b200.so")):
        ge.build()
    return ge.load_package()


def test_library_exports_every_declared_symbol(pkg):
    capi = pkg._capi
    names = capi.header_symbols()
    assert len(names) >= 40
    l = capi.lib()
    for n in names:
        assert hasattr(l, n), f"{n} declared in include/nts_b200.h but not exported"
    assert set(names) == set(capi._SIGS), "ctypes signatures out of sync with the header"
    assert l.nb_abi_version() == 2


def test_library_is_sm100a_only(pkg):
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", pkg._capi.LIB_PATH], capture_output=True, text=True).stdout
    archs = {line.split(".")[-2] for line in out.splitlines() if line.strip().endswith(".cubin")}
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.NtsError) as e:
        pkg.Cuda_Stream(0)
    assert "CUDA" in str(e.value) or "cuda" in str(e.value)


def test_product_never_imports_oracle():
    for root, _, files in os.walk(ge.PKG_DIR):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                text = open(os.path.join(root, f)).read()
                assert "oracle" not in text.replace("oracle/", "").lower() or "import oracle" not in text, f
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f


def test_host_csc_build_matches_oracle(pkg):
    import oracle
    rng = np.random.default_rng(3)
    V = 500
    pairs = rng.integers(0, V, size=(4000, 2)).astype(np.uint32)
    co, ri, ind, outd = pkg.FullyRepGraph.build_csc_host(pairs, V)
    oco, ori = oracle.build_csc(pairs, V)
    oin, oout = oracle.degrees(pairs, V)
    assert np.array_equal(co, oco) and np.array_equal(ri, ori)
    assert np.array_equal(ind, oin) and np.array_equal(outd, oout)


def test_layer_view_struct_layout(pkg):
    assert ctypes.sizeof(pkg._capi.LayerView) == 16 + 14 * 8
