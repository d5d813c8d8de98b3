#!/bin/bash
# same-call sweep: how the next batches' sampler graphs share the SMs with the running aggregation (sampler kernel family, blocks per SM,
# graph-node priority, aggregation grid)
export NB_BENCH_GRAPH_CACHE=/dev/shm/nb_reddit_graph
python -c "import bench; bench.reddit_shaped_graph(1.0)" 2>/dev/null
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs --modes fused,api --timeline 80"
G="--opt sampler_fused=0"
while read -r extra; do
  echo "[$extra]:"
  timeout 300 $B $extra 2> /tmp/err.txt | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("   ms_per_step", round(d["ms_per_step"],5), d["run"]["windows_ms_per_step"], "e2e", round(d["e2e"]["ms_per_step"],5), "sampler alone us", d["sampler_alone_us_per_batch"])'
  grep timeline /tmp/err.txt | sed 's/^/   /' | cut -c1-330
  grep -i "error\|Traceback" /tmp/err.txt | head -3
done <<LIST
$G
$G --opt sampler_blocks_per_sm=1
$G --opt sampler_blocks_per_sm=2
--opt sampler_blocks_per_sm=1
--opt sampler_blocks_per_sm=2
--opt sampler_block_threads=512 --opt sampler_blocks_per_sm=1
$G --opt sampler_capture_priority=0
--opt sampler_tail=0 --opt sampler_block_threads=512 --opt sampler_capture_priority=0
$G --opt sampler_blocks_per_sm=1 --pipeline 6 --sample-streams 3
$G --opt sampler_blocks_per_sm=1 --opt agg_blocks_per_sm=2
$G --opt agg_blocks_per_sm=2
$G --opt sampler_blocks_per_sm=1 --opt agg_blocks_per_sm=4
$G
$G --opt sampler_blocks_per_sm=2
$G --opt sampler_blocks_per_sm=2 --sample-priority 0
$G --opt sampler_blocks_per_sm=1 --sample-priority 0
$G --sample-priority 0
$G --opt sampler_blocks_per_sm=3
$G --opt sampler_blocks_per_sm=4
$G --opt sampler_blocks_per_sm=2 --opt agg_blocks_per_sm=8
$G --opt sampler_blocks_per_sm=2 --opt agg_blocks_per_sm=5
$G --opt sampler_blocks_per_sm=2 --pipeline 3 --api-pipeline 3
--opt sampler_blocks_per_sm=0
--opt sampler_tail=0 --opt sampler_block_threads=512
$G --opt sampler_blocks_per_sm=2
LIST
