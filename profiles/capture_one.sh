#!/bin/bash
# usage: capture_one.sh <kernel regex> <skip> <tag> -- <command...>   (one --set full capture -> raw + source CSV)
set -u
K=$1; S=$2; T=$3; shift 4
O=gpurun_out
ncu --set full --import-source on --clock-control none -k regex:$K -s $S -c 1 -o $O/one_$T -f "$@" > $O/ncu_one_$T.log 2>&1
ncu -i $O/one_$T.ncu-rep --page raw --csv > $O/r1_raw_$T.csv 2>/dev/null
ncu -i $O/one_$T.ncu-rep --page source --print-source cuda,sass --csv > $O/r1_src_$T.csv 2>/dev/null
rm -f $O/one_$T.ncu-rep
ls -la $O/r1_raw_$T.csv $O/r1_src_$T.csv
