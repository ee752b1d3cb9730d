#!/bin/bash
# Dynamic instruction counts of variant builds of the library (run on the GPU box).
# usage: ab_counts.sh <fixture> <shape> lib1.so lib2.so ...
fx=$1; sh=$2; shift 2
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__sass_inst_executed_op_local_ld.sum,smsp__sass_inst_executed_op_local_st.sum,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,gpu__time_duration.sum,smsp__inst_executed_op_branch.sum
for l in "$@"; do
  echo "== $l shape $sh"
  QR_B200_LIB=$PWD/$l QR_B200_SHAPE=$sh ncu --metrics $M --clock-control none -k regex:qr_render -s 2 -c 1 python tools/prof_run.py $fx 3 2>&1 | grep -E "^\s+(smsp|sm__|l1tex|gpu__)" | sed 's/  */ /g'
done
