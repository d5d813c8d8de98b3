#!/bin/bash
# one 1-GPU call: parity tests, step timeline, keep-threshold sweep, default line
export NB_BENCH_GRAPH_CACHE=/dev/shm/nb_reddit_graph
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_reference_trainer.py --deselect tests/test_gpu_multi.py 2>&1 | tail -8 > gpurun_out/r2n_tests.log
python -c "import bench; bench.reddit_shaped_graph(1.0)" 2>/dev/null
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs"
$B --modes fused --timeline 80 2> gpurun_out/r2n_timeline.err | tail -1 > gpurun_out/r2n_timeline.json
for k in 2 3 4 2 3 4; do
  echo -n "keep_min=$k : "; $B --modes fused --opt gather_keep_min_uses=$k 2>/dev/null | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["run"]["windows_ms_per_step"], d["roofline"]["kernels"])'
done > gpurun_out/r2n_keep_sweep.txt 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
du -sh gpurun_out
