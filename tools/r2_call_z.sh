#!/bin/bash
export NB_BENCH_GRAPH_CACHE=/dev/shm/nb_reddit_graph
python -c "import bench; bench.reddit_shaped_graph(1.0)" 2>/dev/null
timeout 600 python -m pytest -m gpu -q -x tests/test_gpu_multi.py 2>&1 | tail -3
run() { N=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
        bench.py --gpus $N --steps 20 --warmup 5 "$@" 2>/tmp/err_$N.txt | tail -1; grep timeline /tmp/err_$N.txt | cut -c1-400 | sed 's/^/      /' >&2; }
fmt='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("value_ms", round(d["ms_per_step"],4), d["run"]["windows_ms_per_step"], "check", (d.get("exchange_check") or "")[:14], "wait_us", d.get("exchange_wait_us"))'
COMMON="--no-cpu-baseline --no-other-configs --modes fused --timeline 60"
for EX in split one split one; do
  echo -n "N=2 exchange=$EX : "
  run 2 $COMMON --exchange $EX 2>/tmp/tl.txt | python -c "$fmt"; cat /tmp/tl.txt
done
