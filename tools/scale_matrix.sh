#!/bin/bash
# One call on an 8-GPU box: where to put the dense-gradient exchange, measured at N=4 and N=8 side by side, then the default line.
# usage: bash tools/scale_matrix.sh > gpurun_out/scale_matrix.txt
export NB_BENCH_GRAPH_CACHE=/dev/shm/nb_reddit_graph
python -c "import bench; bench.reddit_shaped_graph(1.0)" 2>/dev/null
run() { N=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
        bench.py --gpus $N --steps 20 --warmup 5 "$@" 2>/dev/null | tail -1; }
fmt='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("value_ms", round(d["ms_per_step"],4), d["run"]["windows_ms_per_step"], "e2e_ms", round(d["e2e"]["ms_per_step"],4), "check", (d.get("exchange_check") or "")[:14], "wait_us", d.get("exchange_wait_us"))'
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs --modes fused,api 2>/dev/null | tail -1 | python -c "$fmt" | sed 's/^/N=1 : /'
for cfg in "4 split" "8 split" "8 one" "8 nccl"; do
  set -- $cfg; N=$1; EX=$2; shift 2
  echo -n "N=$N exchange=$EX $* : "
  run $N --no-cpu-baseline --no-other-configs --modes fused,api --exchange $EX "$@" | python -c "$fmt"
done
# the default line (what the driver runs), with the other configs: GAT strong scaling, products, full-size papers100M sharded over the 8 GPUs
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29911 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8_default.json 2> gpurun_out/r2_bench_n8_default.err
tail -c 600 gpurun_out/r2_bench_n8_default.err
