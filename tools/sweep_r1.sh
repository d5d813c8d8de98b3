# in-situ comparison of gather variants (TMA bulk rows vs register path) and sampler-stream priority on the headline workload
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/gather_bench.py 2>&1 | head -8
B="python bench.py --steps 400 --warmup 10 --no-cpu-baseline"
for cfg in "NB_GATHER_VARIANT=1 PRI=0" "NB_GATHER_VARIANT=1 PRI=-1" "NB_GATHER_VARIANT=0 PRI=-1" "NB_GATHER_VARIANT=0 PRI=0"; do
  echo "== $cfg"; env $cfg bash -c "$B --sample-priority \$PRI" 2>/dev/null | python tools/bench_line.py
done
