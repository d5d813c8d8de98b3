"""End-to-end drop-in check on the GPU: oracle/_ref/nts_b200 is the reference's OWN trainer (toolkits/main.cpp + core/ + comm/,
compiled unchanged against sample-based-gnn_b200/host/cuda/ntsCUDA.hpp and linked to libnts_b200.so instead of the reference's
CUDA library; `make -C oracle nts`). It must train cora through every sampled toolkit with the kernels of this repo underneath.
The binary and its inputs are staged in the build container (oracle/stage_trainer.py) and travel to the GPU box."""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
BIN = os.path.join(REFDIR, "nts_b200")

TOOLKITS = ["GCNSAMPLEALLGPU", "GCNSAMPLEGPU", "GSSAMPLEALLGPU", "GATSAMPLEALLGPU", "GSSAMPLECACHE", "GCNSAMPLEPDCACHE", "GSSAMPLEPDCACHE",
            "GATSAMPLEPDCACHE"]
# CACHE:1 (the shipped gcn_reddit_sample.cfg's setting): the feature load goes through FastSampler::load_feature_gpu_cache ->
# zero_copy_feature_move_gpu_cache + gather_feature_from_gpu_cache (core/ntsFastSampler.hpp:263-317)
CACHED = ["GCNSAMPLEPDCACHE_cache1", "GSSAMPLEPDCACHE_cache1", "GATSAMPLEPDCACHE_cache1"]
# one thread per GPU + the adaptor's NCCL_Communicator; GPU_NUM:1 runs everywhere, GPU_NUM:2 where the box has two GPUs
MULTI = ["GCNSAMPLEALLMULTI", "GATSAMPLEALLMULTI", "GCNSAMPLEPCMULTI", "GSSAMPLEPCMULTI", "GATSAMPLEPCMULTI",
         "GCNSAMPLEPCMULTI_cache1", "GSSAMPLEPCMULTI_cache1", "GATSAMPLEPCMULTI_cache1"]
CASES = TOOLKITS + CACHED + [f"{m}_g{g}" for m in MULTI for g in (1, 2)]
# GAT_SAMPLE_PC_MULTI with CACHE:1 reads `outmost_vertex` / `dev_cache_feature`, which only determine_cache_node_idx() allocates and
# which that toolkit never calls (toolkits/GAT_SAMPLE_PC_MULTI.hpp:845-907 vs its run(): no call site): a reference bug on a path it
# cannot have exercised. Expected to fail in the reference's own host code, before any kernel of this library is involved.
REFERENCE_BUGS = {"GATSAMPLEPCMULTI_cache1_g1", "GATSAMPLEPCMULTI_cache1_g2"}


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/nts_b200 not built (needs /root/reference at build time)")
@pytest.mark.parametrize("case", CASES)
def test_reference_trainer_runs_on_libnts_b200(case):
    if case.endswith("_g2") and _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    if case in REFERENCE_BUGS:
        pytest.xfail("reference bug: the toolkit never allocates the buffers its CACHE:1 branch uses (see REFERENCE_BUGS)")
    import glob
    for f in glob.glob(os.path.join(REFDIR, "data", "*pre_sample*.bin")):   # hot-vertex lists a previous toolkit left behind
        os.remove(f)
    r = subprocess.run([BIN, f"cfg_{case}.cfg"], cwd=REFDIR, capture_output=True, text=True, timeout=300)
    out = r.stdout + r.stderr
    if os.environ.get("NB_TRAINER_LOGS"):              # keep every case's full output (diagnostics on the GPU box)
        os.makedirs(os.environ["NB_TRAINER_LOGS"], exist_ok=True)
        open(os.path.join(os.environ["NB_TRAINER_LOGS"], f"trainer_{case}.log"), "w").write(f"rc={r.returncode}\n" + out)
    assert r.returncode == 0, out[-3000:]
    assert "is not provided by libnts_b200" not in out, out[-2000:]
    accs = [float(m.group(1)) for m in re.finditer(r"Train Acc: ([0-9.]+)", out)]
    losses = [float(m.group(1)) for m in re.finditer(r"Epoch\[\d+\]:Times\[[^\]]*\]:loss\s+([0-9.eE+-]+)", out)]
    assert len(accs) >= 4, out[-2000:]
    if "MULTI" not in case:                           # the *_MULTI toolkits print the epoch time without the loss
        assert len(losses) >= 4, out[-2000:]
    # the *_PD_CACHE and *_PC_MULTI toolkits train on bounded-stale hot embeddings: they plateau around 0.55-0.72 on cora within 10 epochs and move by
    # several points from run to run (clock-seeded shuffles): they must clearly learn (chance = 0.14), the plain toolkits must reach 0.70
    stale = "PDCACHE" in case or "PCMULTI" in case
    floor = 0.50 if stale else 0.70
    assert max(accs[-2:]) >= floor, accs              # cora, 5 epochs (10 for the toolkits that train on bounded-stale hot embeddings;
                                                      # the reference's own log reaches 0.93 after 10 epochs of the plain toolkits)
    assert not losses or (min(losses) if stale else losses[-1]) < losses[0], losses   # stale-embedding toolkits: the loss is noisy epoch to epoch
