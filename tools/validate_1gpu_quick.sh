export NB_BENCH_GRAPH_CACHE=/dev/shm/nb_reddit_graph
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/r2c_smoke.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r2c_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2c_ref.json 2> gpurun_out/r2c_ref.err
