#!/bin/bash
# Round-1 profile capture (run under gpurun on one B200; outputs small CSVs only).
# 1. plain runs must exit 0 first; 2. launch lists (gpu__time_duration per launch, compare SHARES only);
# 3. one `--set full` capture per kernel, written straight to CSV (no .ncu-rep kept: they exceed the 64 MiB pull limit).
set -u
O=gpurun_out
python bench.py --steps 2 --warmup 3 --chain-only > $O/plain_chain.log 2>&1 || exit 1
python bench.py --ops-only > $O/plain_ops.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r1_launches_chain.csv \
    python bench.py --steps 2 --warmup 3 --chain-only > $O/ncu_a.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r1_launches_ops.csv \
    python bench.py --ops-only > $O/ncu_c.log 2>&1
for k in nms_rank_kernel nms_mask_kernel nms_reduce_kernel soma_binarize_kernel largest_cc_fill_kernel largest_cc_kernel paste_bin_kernel paste_labels_kernel; do
  ncu --set full --clock-control none -k regex:$k -s 3 -c 1 --csv --page raw --log-file $O/r1_raw_$k.csv \
      python bench.py --steps 2 --warmup 3 --chain-only > $O/ncu_$k.log 2>&1
done
for k in roialign3d_fwd_kernel roialign3d_bwd_kernel peaks_scan3_kernel peaks_refine_kernel peaks_filter_kernel peaks_emit_kernel iou3d_kernel gauss3d_kernel median3d_kernel mask_joint_hist_kernel rle_walk_kernel rle_fill_kernel gp_hist_kernel gp_rank_decode_kernel zs_apply_kernel; do
  ncu --set full --clock-control none -k regex:$k -s 3 -c 1 --csv --page raw --log-file $O/r1_raw_$k.csv \
      python bench.py --ops-only > $O/ncu_$k.log 2>&1
done
ls -la $O | tail -30
