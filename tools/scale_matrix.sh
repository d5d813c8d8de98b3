#!/bin/bash
# One call on an N-GPU box: weak scaling of the headline step (default settings), same box, same call; optional exchange A/B.
# usage: bash tools/scale_matrix.sh [full] > gpurun_out/scale_matrix.txt        (full: also N=4 and the exchange as split / nccl)
export NB_BENCH_GRAPH_CACHE=/dev/shm/nb_reddit_graph
python -c "import bench; bench.reddit_shaped_graph(1.0)" 2>/dev/null
run() { N=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
        bench.py --gpus $N --steps 20 --warmup 5 "$@" 2>/tmp/err_$N.txt | tail -1; grep timeline /tmp/err_$N.txt | cut -c1-420 | sed 's/^/      /' >&2; }
fmt='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("value_ms", round(d["ms_per_step"],4), d["run"]["windows_ms_per_step"], "check", (d.get("exchange_check") or "")[:14], "wait_us", d.get("exchange_wait_us"), "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])'
COMMON="--no-cpu-baseline --no-other-configs --modes fused --timeline 60"
python bench.py --steps 20 --warmup 5 $COMMON 2>/tmp/err_1.txt | tail -1 | python -c "$fmt" | sed 's/^/N=1 : /'; grep timeline /tmp/err_1.txt | cut -c1-420 | sed 's/^/      /'
CFGS=("8 one" "8 one")
[ "$1" = full ] && CFGS=("8 one" "8 split" "8 nccl" "4 one" "8 one")
for cfg in "${CFGS[@]}"; do
  set -- $cfg; N=$1; EX=$2
  echo -n "N=$N exchange=$EX : "
  run $N $COMMON --exchange $EX 2>/tmp/tl.txt | python -c "$fmt"; cat /tmp/tl.txt
done
