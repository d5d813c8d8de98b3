#!/bin/bash
export NB_BENCH_GRAPH_CACHE=/dev/shm/nb_reddit_graph
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "short_row or aggregate or long_segments or replay or gat" 2>&1 | tail -5 > gpurun_out/r2p_tests.log
python -c "import bench; bench.reddit_shaped_graph(1.0)" 2>/dev/null
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs --modes fused --timeline 80"
for extra in "" "--opt agg_short_rows=0" "--opt agg_deep_small=0" "--opt agg_short_rows=0 --opt agg_deep_small=0" "--pipeline 6 --sample-streams 3" ""; do
  echo "default + [$extra]:"
  $B $extra 2> /tmp/err.txt | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("   ms_per_step", round(d["ms_per_step"],5), d["run"]["windows_ms_per_step"])'
  grep timeline /tmp/err.txt | sed 's/^/   /'
done > gpurun_out/r2p_top_hop_ab.txt 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
du -sh gpurun_out
