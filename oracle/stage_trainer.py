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

ALGORITHMS = ["GCNSAMPLEALLGPU", "GCNSAMPLEGPU", "GSSAMPLEALLGPU", "GATSAMPLEALLGPU", "GSSAMPLECACHE", "GCNSAMPLEPDCACHE",
              "GCNSAMPLEALLMULTI", "GATSAMPLEALLMULTI", "GCNSAMPLESINGLE"]

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
CACHE:0
GPU_NUM:{gpus}
"""


def main():
    data = os.path.join(OUT, "data")
    os.makedirs(data, exist_ok=True)
    for f in ("cora.2708.edge.self", "cora.labeltable", "cora.mask"):
        shutil.copyfile(os.path.join(REF, "data", f), os.path.join(data, f))
    with zipfile.ZipFile(os.path.join(REF, "data", "cora.featuretable.zip")) as z:
        z.extractall(data)
    for alg in ALGORITHMS:
        multi = "MULTI" in alg
        with open(os.path.join(OUT, f"cfg_{alg}.cfg"), "w") as f:
            f.write(CFG.format(alg=alg, batch=1024, epochs=5, pipeline=2 if "CACHE" in alg else 1, gpus=2 if multi else 1))
    print("staged", len(ALGORITHMS), "cfgs under", OUT)


if __name__ == "__main__":
    main()
