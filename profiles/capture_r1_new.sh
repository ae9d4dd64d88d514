#!/bin/bash
# Round-1, second batch: `--set full` captures of the kernels added after capture_r1.sh ran (Mask R-CNN mask paste-back,
# nuclei chain).  Run under gpurun on one B200 AFTER the plain command has exited 0; outputs small CSVs only.
set -u
O=gpurun_out
python bench.py --ops-only > $O/plain_ops2.log 2>&1 || exit 1
for k in segm_resize_paste_kernel nuclei_minmax_kernel nuclei_normalise_kernel nuclei_morph_kernel otsu2d_kernel mask_complement_kernel; do
  ncu --set full --clock-control none -k regex:$k -s 3 -c 1 --csv --page raw --log-file $O/r1_raw_$k.csv \
      python bench.py --ops-only > $O/ncu_$k.log 2>&1
done
ls -la $O/r1_raw_segm* $O/r1_raw_nuclei* $O/r1_raw_otsu2d* $O/r1_raw_mask_complement*
