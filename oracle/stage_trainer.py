"""Stages what oracle/_ref/nts_b200 (the reference's own trainer, compiled unchanged against the adaptor header and linked to
libnts_b200.so -- `make -C oracle nts`) needs to run on the GPU box: the cora fixture files of the reference's data/ directory and
one .cfg per toolkit. TEST INFRASTRUCTURE ONLY; everything goes under oracle/_ref/ (git-ignored, travels with gpurun).

    python oracle/stage_trainer.py        # needs /root/reference (build container)
"""
import os
import shutil
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("NTS_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")

# every toolkit toolkits/main.cpp dispatches to (:67-186)
ALGORITHMS = ["GCNSAMPLEALLGPU", "GCNSAMPLEGPU", "GSSAMPLEALLGPU", "GATSAMPLEALLGPU", "GSSAMPLECACHE", "GCNSAMPLEPDCACHE",
              "GSSAMPLEPDCACHE", "GATSAMPLEPDCACHE", "GCNSAMPLEALLMULTI", "GATSAMPLEALLMULTI", "GCNSAMPLEPCMULTI", "GSSAMPLEPCMULTI",
              "GATSAMPLEPCMULTI", "GCNSAMPLESINGLE"]
# toolkits whose feature load goes through FastSampler::load_feature_gpu_cache when the cfg says CACHE:1 (graph->config->cacheflag,
# e.g. toolkits/GCN_SAMPLE_PD_CACHE.hpp:535-545): staged a second time with CACHE:1, the setting of the shipped gcn_reddit_sample.cfg
CACHED_FEATURE_TOOLKITS = ["GCNSAMPLEPDCACHE", "GSSAMPLEPDCACHE", "GATSAMPLEPDCACHE", "GCNSAMPLEPCMULTI", "GSSAMPLEPCMULTI", "GATSAMPLEPCMULTI"]

CFG = """ALGORITHM:{alg}
VERTICES:2708
LAYERS:1433-256-7
FANOUT:25-10
BATCH_SIZE:{batch}
EPOCHS:{epochs}
EDGE_FILE:./data/cora.2708.edge.self
FEATURE_FILE:./data/cora.featuretable
LABEL_FILE:./data/cora.labeltable
MASK_FILE:./data/cora.mask
LEARN_RATE:0.01
WEIGHT_DECAY:0.0001
DECAY_RATE:0.97
DECAY_EPOCH:100
DROP_RATE:0.5
PIPELINE_NUM:{pipeline}
CACHE_RATE:0.2
FEATURE_CACHE_RATE:0.2
UP_DEGREE:0
PROC_OVERLAP:0
PROC_LOCAL:0
PROC_CUDA:0
PROC_REP:0
LOCK_FREE:1
PUSHDOWN:0
CACHE:{cache}
GPU_NUM:{gpus}
"""


def main():
    data = os.path.join(OUT, "data")
    os.makedirs(data, exist_ok=True)
    for f in ("cora.2708.edge.self", "cora.labeltable", "cora.mask"):
        shutil.copyfile(os.path.join(REF, "data", f), os.path.join(data, f))
    with zipfile.ZipFile(os.path.join(REF, "data", "cora.featuretable.zip")) as z:
        z.extractall(data)
    n = 0
    for alg in ALGORITHMS:
        multi = "MULTI" in alg
        pipeline = 2 if ("CACHE" in alg or "PCMULTI" in alg) else 1
        if alg == "GATSAMPLEPDCACHE":
            # its edge-NN lambdas read the shared graph->rtminfo->curr_layer (toolkits/GAT_SAMPLE_PD_CACHE.hpp:428-455) while the other
            # pipeline thread advances it: with two pipeline threads the reference's own autograd bookkeeping asserts
            # (core/ntsContext.hpp:491). GAT_SAMPLE_PC_MULTI.hpp fixed this with `int layer = i`. One pipeline slot avoids the race.
            pipeline = 1
        epochs = 10 if alg in ("GSSAMPLEPDCACHE", "GATSAMPLEPDCACHE", "GSSAMPLEPCMULTI", "GATSAMPLEPCMULTI") else 5   # stale hot embeddings converge slower
        for cache in ([0, 1] if alg in CACHED_FEATURE_TOOLKITS else [0]):
            for gpus in ([1, 2] if multi else [1]):       # the *_MULTI toolkits also run on one device (GPU_NUM:1)
                name = f"cfg_{alg}" + ("_cache1" if cache else "") + (f"_g{gpus}" if multi else "") + ".cfg"
                with open(os.path.join(OUT, name), "w") as f:
                    f.write(CFG.format(alg=alg, batch=1024, epochs=epochs, pipeline=pipeline, gpus=gpus, cache=cache))
                n += 1
    print("staged", n, "cfgs under", OUT)


if __name__ == "__main__":
    main()
