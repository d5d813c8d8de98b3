#!/bin/bash
export NB_BENCH_GRAPH_CACHE=/dev/shm/nb_reddit_graph
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2q_tests.log
python -c "import bench; bench.reddit_shaped_graph(1.0)" 2>/dev/null
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs --modes fused,api --timeline 80"
for extra in "" "--opt sampler_block_threads=512" "--opt sampler_tail=0 --opt sampler_block_threads=512" "--opt sampler_fused=0" "--sample-streams 1 --pipeline 2 --api-pipeline 2" "--sample-streams 1 --pipeline 2 --api-pipeline 2 --opt sampler_tail=0 --opt sampler_block_threads=512" ""; do
  echo "default + [$extra]:"
  timeout 300 $B $extra 2> /tmp/err.txt | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("   ms_per_step", round(d["ms_per_step"],5), d["run"]["windows_ms_per_step"], "e2e", round(d["e2e"]["ms_per_step"],5), "host wait", d["e2e"]["host_blocked_in_sampler_wait_ms_per_step"])'
  grep timeline /tmp/err.txt | sed 's/^/   /'
  grep -i "error\|Traceback" /tmp/err.txt | head -3
done > gpurun_out/r2q_sampler_ab.txt 2>&1
du -sh gpurun_out
