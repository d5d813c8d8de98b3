#!/bin/bash
# same-call A/B: the weight-free bottom hop on its own stream (batch i+1's Y1 = A X0 beside batch i's top hop) vs everything in the training stream
export NB_BENCH_GRAPH_CACHE=/dev/shm/nb_reddit_graph
python -c "import bench; bench.reddit_shaped_graph(1.0)" 2>/dev/null
B="python bench.py --steps 20 --warmup 5 --no-other-configs --cpu-batches 4 --timeline 80"
while read -r extra; do
  echo "[$extra]:"
  timeout 400 $B $extra 2> /tmp/err.txt | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("   ms_per_step", round(d["ms_per_step"],5), d["run"]["windows_ms_per_step"], "e2e", round(d["e2e"]["ms_per_step"],5), "cpp", (d["e2e"].get("cpp_host") or {}).get("ms_per_step"), "materialized", round(d["materialized_x0"]["ms_per_step"],5), "kernels", {k: v["ms"] for k, v in d["roofline"]["kernels"].items()})'
  grep timeline /tmp/err.txt | sed 's/^/   /' | cut -c1-420
  grep -i "error\|Traceback" /tmp/err.txt | head -3
done <<LIST

--agg-stream 0
--pipeline 6 --sample-streams 3
--opt sampler_blocks_per_sm=1

LIST
