#!/bin/bash
# ncu evidence for profiles/: every command first runs plain (must exit 0), then under ncu with --clock-control none.
# (1) launch list of the headline bench (serial pipeline so that shares are readable), (2) --set full of every hot kernel.
set -u
OUT=gpurun_out
M="launch__grid_size,launch__block_size,launch__registers_per_thread,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
BENCH="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --pipeline 1"
$BENCH > $OUT/ncu_plain_bench.log 2>&1 || { echo "plain bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r1b_ncu_launches_bench_pipeline1.csv $BENCH > /dev/null 2>&1
full() {  # name, kernel regex, skip, count, command...
  local name=$1 re=$2 skip=$3 cnt=$4; shift 4
  "$@" > $OUT/ncu_plain_$name.log 2>&1 || { echo "plain $name failed"; return; }
  ncu --set full --clock-control none --import-source on --kernel-name "regex:$re" --launch-skip $skip --launch-count $cnt -f -o $OUT/r1b_full_$name "$@" > $OUT/ncu_$name.log 2>&1
  ncu -i $OUT/r1b_full_$name.ncu-rep --page raw --csv --metrics $M > $OUT/r1b_full_$name.csv 2>/dev/null
}
full bench "k_gather_rows_tma|k_segment_reduce|k_sample|k_relabel|k_scan|k_csr" 60 14 $BENCH
full gat "k_gat_|k_segment_reduce" 40 8 python tools/gat_bench.py
full narrow "k_gather_rows_narrow" 10 2 python tools/config_bench.py --config products --steps 12 --warmup 2
full ingest "k_rs_|k_split_pairs" 0 7 python tools/ingest_bench.py
ls -la $OUT/r1b_* | awk '{print $5, $9}'
