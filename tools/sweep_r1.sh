OLD=$PWD/sample-based-gnn_b200/lib/libnts_b200_875e019.so
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
echo "== agg microbench old"; NB_LIB_PATH=$OLD python tools/agg_bench.py 2>&1 | tail -4
echo "== agg microbench new"; python tools/agg_bench.py 2>&1 | tail -4
B="python bench.py --steps 600 --warmup 10 --no-cpu-baseline"
echo "old:"; NB_LIB_PATH=$OLD $B 2>/dev/null | python tools/bench_line.py
echo "new:"; $B 2>/dev/null | python tools/bench_line.py
echo "== gat old"; NB_LIB_PATH=$OLD python tools/gat_bench.py 2>&1 | grep "fused bwd"
echo "== gat new"; python tools/gat_bench.py 2>&1 | grep "fused bwd"
