#!/bin/bash
# ncu captures of the prefilter kernels (run under gpurun, one B200): raw metrics + per-line source page as CSV.
set -u
O=gpurun_out
python profiles/prof_prefilter.py > $O/plain_prefilter.log 2>&1 || exit 1
for k in gauss3d_kernel median3d_kernel; do
  ncu --set full --import-source on --clock-control none -k regex:$k -s 2 -c 1 -o $O/pre_$k -f python profiles/prof_prefilter.py > $O/ncu_$k.log 2>&1
  ncu -i $O/pre_$k.ncu-rep --page raw --csv > $O/r1_raw_$k.csv 2>/dev/null
  ncu -i $O/pre_$k.ncu-rep --page source --print-source cuda,sass --csv > $O/r1_src_$k.csv 2>/dev/null
  rm -f $O/pre_$k.ncu-rep
done
ls -la $O | tail -8
