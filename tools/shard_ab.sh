#!/bin/bash
# N GPUs: the sharded-table gather through the register path vs TMA bulk copies (rows of peer shards over NVLink), same call
N=${1:-2}
for opt in "" "table_gather_tma=1" "" "table_gather_tma=1"; do   # SHARD_ROWS=80000 for a mini-batch sized gather
  echo "== [$opt]"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) tools/shard_bench.py $opt 2>&1 | grep -E "SHARD|option|Error|error|assert" | cut -c1-260
done
