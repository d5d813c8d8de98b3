#!/bin/bash
# Same-call sweep: how many resident aggregation blocks per SM leave room for the sampler's kernels to run under the aggregation.
for FUSED in 1 0; do for BPS in 8 7 6 5; do
  echo -n "sampler_fused=$FUSED agg_blocks_per_sm=$BPS : "
  python bench.py --steps 50 --warmup 10 --windows 3 --no-cpu-baseline --no-other-configs --modes fused --opt sampler_fused=$FUSED --opt agg_blocks_per_sm=$BPS 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['run']['windows_ms_per_step'], d['roofline']['kernels'][d['roofline']['kernel']]['ms'])"
done; done
