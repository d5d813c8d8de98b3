#!/bin/bash
# ncu evidence for profiles/ (round 2): every command first runs plain (must exit 0), then under ncu with --clock-control none.
# (1) launch list of the headline bench (serial pipeline so that shares are readable), (2) --set full of every hot kernel.
# usage (GPU box, one GPU): bash tools/ncu_capture.sh ; copy gpurun_out/r2_* into profiles/
set -u
OUT=gpurun_out
M="launch__grid_size,launch__block_size,launch__registers_per_thread,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
BENCH="python bench.py --steps 4 --warmup 3 --windows 1 --no-cpu-baseline --no-other-configs --pipeline 1 --modes fused,materialized"
$BENCH > $OUT/ncu_plain_bench.log 2>&1 || { echo "plain bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/r2b_ncu_launches_bench_pipeline1.csv $BENCH > /dev/null 2>&1
full() {  # name, kernel regex, skip, count, command...
  local name=$1 re=$2 skip=$3 cnt=$4; shift 4
  "$@" > $OUT/ncu_plain_$name.log 2>&1 || { echo "plain $name failed"; return; }
  ncu --set full --clock-control none --import-source on --kernel-name "regex:$re" --launch-skip $skip --launch-count $cnt -f -o $OUT/r2b_full_$name "$@" > $OUT/ncu_$name.log 2>&1
  ncu -i $OUT/r2b_full_$name.ncu-rep --page raw --csv --metrics $M > $OUT/r2b_full_$name.csv 2>/dev/null
  # the reports themselves are tens of MB each (gpurun_out/ travels back only below 64 MiB): keep the summaries, and for the dominant
  # kernel the per-instruction source page
  if [ "$name" = bench ]; then ncu -i $OUT/r2b_full_$name.ncu-rep --page source --csv --kernel-name "regex:k_segment_reduce" 2>/dev/null | head -400 > $OUT/r2b_source_segment_reduce.csv; fi
  rm -f $OUT/r2b_full_$name.ncu-rep
}
# the device-resident (fused) arm runs first: skip its warm-up launches, then take two steps' worth of every kernel
full bench "k_segment_reduce|k_sample|k_relabel|k_scan|k_csr|k_pack_gather" 60 30 $BENCH
full bench_noh "k_segment_reduce" 12 4 $BENCH --no-l2-hints
full gat "k_gat_|k_segment_reduce" 40 10 python tools/gat_bench.py
full papers "k_sample|k_scan|k_relabel|k_csr|k_pack" 22 11 python tools/papers_sampler_prof.py
ls -la $OUT/r2b_* | awk '{print $5, $9}'
