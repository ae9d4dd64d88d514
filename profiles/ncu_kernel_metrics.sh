#!/bin/bash
# usage: ncu_kernel_metrics.sh <kernel regex> <count> <tag> -- <command...>: a few whole-kernel metrics per launch -> gpurun_out/km_<tag>.csv
K=$1; N=$2; T=$3; shift 4
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:$K -c $N --csv --log-file gpurun_out/km_$T.csv "$@" > gpurun_out/km_$T.log 2>&1
