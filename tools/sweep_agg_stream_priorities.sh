export NB_BENCH_GRAPH_CACHE=/dev/shm/nb_reddit_graph
python -c "import bench; bench.reddit_shaped_graph(1.0)" 2>/dev/null
B="python bench.py --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline --modes fused,api --timeline 80"
while read -r extra; do
  echo "[$extra]:"
  timeout 400 $B $extra 2> /tmp/err.txt | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("   ms_per_step", round(d["ms_per_step"],5), d["run"]["windows_ms_per_step"], "e2e", round(d["e2e"]["ms_per_step"],5))'
  grep timeline /tmp/err.txt | sed 's/^/   /' | cut -c1-420
  grep -i "error\|Traceback" /tmp/err.txt | head -3
done <<LIST
--train-priority -1
--train-priority -2
--train-priority -2 --sample-priority -2
--train-priority -1 --opt agg_deep_small=0
--train-priority -2 --pipeline 6 --sample-streams 3
--train-priority -1 --sample-priority 0

LIST
