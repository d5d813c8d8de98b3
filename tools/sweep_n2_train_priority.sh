#!/bin/bash
# N=2, same call: with the bottom hop on its own stream the exchange sits on the weight-dependent chain; does a higher priority for that chain hide it?
export NB_BENCH_GRAPH_CACHE=/dev/shm/nb_reddit_graph
python -c "import bench; bench.reddit_shaped_graph(1.0)" 2>/dev/null
run() { N=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
        bench.py --gpus $N --steps 20 --warmup 5 "$@" 2>/tmp/err_$N.txt | tail -1; grep timeline /tmp/err_$N.txt | cut -c1-420 | sed 's/^/      /' >&2; }
fmt='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("value_ms", round(d["ms_per_step"],4), d["run"]["windows_ms_per_step"], "check", (d.get("exchange_check") or "")[:14], "wait_us", d.get("exchange_wait_us"))'
COMMON="--no-cpu-baseline --no-other-configs --modes fused --timeline 60"
N=${1:-2}
while read -r extra; do
  echo -n "N=$N [$extra] : "
  run $N $COMMON $extra 2>/tmp/tl.txt | python -c "$fmt"; cat /tmp/tl.txt
done <<LIST

--train-priority -1
--train-priority -2
--train-priority -1 --exchange split
--agg-stream 0
--train-priority -1
LIST
