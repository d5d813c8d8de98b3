#!/bin/bash
# Same-call sweep of the knobs that decide how much of the sampler hides behind the aggregation (headline arm only).
# usage (GPU box): bash tools/sweep_pipeline.sh > gpurun_out/sweep_pipeline.txt
for P in 2 3 4; do for PRI in 0 -1; do for OPT in "" "--opt agg_blocks_per_sm=7" "--opt agg_persistent=0"; do
  echo -n "pipeline=$P sample_priority=$PRI $OPT : "
  python bench.py --steps 50 --warmup 10 --windows 3 --no-cpu-baseline --no-other-configs --modes fused --pipeline $P --sample-priority $PRI $OPT 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['run']['windows_ms_per_step'], d['roofline']['kernels'][d['roofline']['kernel']]['ms'])"
done; done; done
