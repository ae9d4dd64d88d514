#!/bin/bash
# Round 2: `ncu --set full` capture of the three RoIAlign3D kernels on BASELINE config 4 (512 RoIs x 256 ch x 7^3, fp32):
# forward fast kernel, backward per-RoI kernel, backward gather kernel (one launch each, after two warm-up launches).
# Run under gpurun after the plain command exited 0.
set -u
O=gpurun_out
python profiles/time_roialign.py > $O/r2_roialign_plain.log 2>&1 || { tail -3 $O/r2_roialign_plain.log; exit 1; }
: > $O/r2_raw_roialign.csv
for k in roialign3d_fwd_fast_kernel roialign3d_bwd_roi_kernel roialign3d_bwd_gather_kernel; do
  ncu --set full --clock-control none -k regex:$k -s 2 -c 1 -o $O/r2_$k -f python profiles/time_roialign.py > $O/r2_roialign_ncu.log 2>&1
  ncu -i $O/r2_$k.ncu-rep --page raw --csv >> $O/r2_raw_roialign.csv 2>/dev/null
  rm -f $O/r2_$k.ncu-rep
done
tail -1 $O/r2_roialign_plain.log
