export NB_BENCH_GRAPH_CACHE=/dev/shm/nb_reddit_graph
python -c "import bench; bench.reddit_shaped_graph(1.0)" 2>/dev/null
python -m pytest -m gpu -q -x tests/test_gpu_multi.py 2>&1 | tail -5
run() { N=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
        bench.py --gpus $N --steps 20 --warmup 5 "$@" 2>/dev/null | tail -1; }
fmt='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("value_ms", round(d["ms_per_step"],4), d["run"]["windows_ms_per_step"], "e2e_ms", round(d["e2e"]["ms_per_step"],4), "check", (d.get("exchange_check") or "")[:14], "wait_us", d.get("exchange_wait_us"))'
for EX in split split-inline one split split-inline one; do
  echo -n "N=2 exchange=$EX : "
  run 2 --no-cpu-baseline --no-other-configs --modes fused --exchange $EX | python -c "$fmt"
done
