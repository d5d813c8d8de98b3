#!/bin/bash
# same-call sweep: pipeline slots x sampling streams (is the sampler graph of the next batch on the step's critical path?)
export NB_BENCH_GRAPH_CACHE=/dev/shm/nb_reddit_graph
python -c "import bench; bench.reddit_shaped_graph(1.0)" 2>/dev/null
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs --modes fused --timeline 80"
for cfg in "2 1" "3 1" "3 2" "4 2" "3 3" "6 3" "2 1" "4 2"; do
  set -- $cfg
  echo "pipeline=$1 sample_streams=$2 $EXTRA:"
  $B --pipeline $1 --sample-streams $2 $EXTRA 2> /tmp/err.txt | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("   ms_per_step", round(d["ms_per_step"],5), d["run"]["windows_ms_per_step"])'
  grep timeline /tmp/err.txt | sed 's/^/   /'
done
