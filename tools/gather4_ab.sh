timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gather4 or tma_gather" 2>&1 | tail -8
timeout 600 python tools/gather_bench.py 2>&1 | tail -40
