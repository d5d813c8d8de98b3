#!/bin/bash
export NB_BENCH_GRAPH_CACHE=/dev/shm/nb_reddit_graph
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r2w_tests.log
python tools/gat_bench.py > gpurun_out/r2w_gat_bench.txt 2>&1
python tools/gat_bench.py agg_short_rows=0 >> gpurun_out/r2w_gat_bench.txt 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err
