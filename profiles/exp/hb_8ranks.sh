#!/bin/bash
# Eight copies of hb_scale.py side by side, one per GPU, 8 volumes each (the per-rank share of the 64-volume step), the same
# sequence of host_batch_mode values in every copy, `secs` seconds per mode so that the copies overlap: loads the host like the
# 8-rank bench does and compares transfer schemes under that load.  usage: hb_8ranks.sh modes secs
modes=${1:-71,7,71,7}; secs=${2:-2}
export LOCAL_WORLD_SIZE=8
for i in 0 1 2 3 4 5 6 7; do
  CUDA_VISIBLE_DEVICES=$i python profiles/exp/hb_scale.py 8 $modes $secs > gpurun_out/hb8_$i.log 2>&1 &
done
wait
for i in 0 1 2 3 4 5 6 7; do tail -n +1 gpurun_out/hb8_$i.log | grep "mode=" | sed "s/^/gpu$i /"; done | sort -k2 -n
