#!/bin/bash
# Round 2: launch list of the timed chain (bench.py --chain-only, launched kernel by kernel so that ncu sees the
# nodes) and one `--set full` capture of one whole step (every kernel of this library, one launch each).
# Run under gpurun on one B200 AFTER the plain command has exited 0; only small CSVs come back.
#   usage: bash profiles/capture_r2.sh [tag]
set -u
O=gpurun_out
T=${1:-r2}
CMD="python bench.py --chain-only --no-graph --steps 2 --warmup 1 --min-timed-ms 0 --volumes 64"
$CMD > $O/${T}_plain_chain.log 2>&1 || { tail -5 $O/${T}_plain_chain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_chain.csv \
    $CMD > $O/${T}_ncu_launches.log 2>&1
# one full step = the 13 launches of this library after the warm-up step (names of every chain kernel)
K='regex:nms_|iota_|soma_binarize|largest_cc|paste_|peaks_'
ncu --set full --import-source on --clock-control none -k "$K" -s 13 -c 13 -o $O/${T}_step -f \
    $CMD > $O/${T}_ncu_full.log 2>&1
ncu -i $O/${T}_step.ncu-rep --page raw --csv > $O/${T}_raw_step.csv 2>/dev/null
ls -la $O/${T}_launches_chain.csv $O/${T}_raw_step.csv $O/${T}_step.ncu-rep
