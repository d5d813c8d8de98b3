#!/bin/bash
# one 1-GPU call: smoke, every GPU test (parity, reference trainer), the default bench line, the reference arm, ncu captures
export NB_BENCH_GRAPH_CACHE=/dev/shm/nb_reddit_graph
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/r2v_smoke.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2v_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2v_ref.json 2> gpurun_out/r2v_ref.err
bash tools/ncu_capture.sh > gpurun_out/r2v_ncu.log 2>&1
du -sh gpurun_out
