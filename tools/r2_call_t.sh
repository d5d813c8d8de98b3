#!/bin/bash
export NB_BENCH_GRAPH_CACHE=/dev/shm/nb_reddit_graph
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2t_tests.log
python -c "import bench; bench.reddit_shaped_graph(1.0)" 2>/dev/null
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs --modes fused,api --timeline 80"
while read -r extra; do
  echo "[$extra]:"
  timeout 300 $B $extra 2> /tmp/err.txt | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("   ms_per_step", round(d["ms_per_step"],5), d["run"]["windows_ms_per_step"], "e2e", round(d["e2e"]["ms_per_step"],5), "sampler alone us", d["sampler_alone_us_per_batch"], "launches", d["gpu_launches"])'
  grep timeline /tmp/err.txt | sed 's/^/   /' | cut -c1-330
  grep -i "error\|Traceback" /tmp/err.txt | head -3
done > gpurun_out/r2t_tail_ab.txt 2>&1 <<LIST

--opt sampler_tail=0
--opt sampler_blocks_per_sm=0
--opt sampler_blocks_per_sm=0 --opt sampler_tail=0
--opt sampler_blocks_per_sm=1
--opt sampler_blocks_per_sm=3

LIST
du -sh gpurun_out
