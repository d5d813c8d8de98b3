#!/bin/bash
# one 2-GPU call: multi-GPU tests, the default bench line at N=2 (all configs), the reference arm under torchrun
export NB_BENCH_GRAPH_CACHE=/dev/shm/nb_reddit_graph
python -c "import bench; bench.reddit_shaped_graph(1.0)" 2>/dev/null
timeout 600 python -m pytest -m gpu -q -x tests/test_gpu_multi.py 2>&1 | tail -3 > gpurun_out/r2x_mgpu_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2x_bench_n2.json 2> gpurun_out/r2x_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2x_ref_n2.json 2> gpurun_out/r2x_ref_n2.err
