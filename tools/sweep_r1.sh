python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/ingest_bench.py 2>&1 | tail -6
