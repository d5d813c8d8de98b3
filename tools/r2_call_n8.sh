#!/bin/bash
export NB_BENCH_GRAPH_CACHE=/dev/shm/nb_reddit_graph
python -c "import bench; bench.reddit_shaped_graph(1.0)" 2>/dev/null
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29911 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2b_bench_n8_default.json 2> gpurun_out/r2b_bench_n8_default.err
tail -c 400 gpurun_out/r2b_bench_n8_default.err
