#!/bin/bash
# ncu --set full of the fused kernel on one size (run on the GPU box through gpurun; one GPU).
# usage: tools/prof_fused.sh <lg> <tag> [extra env assignments...]
LG=${1:-16}; TAG=${2:-fused}; shift 2
for kv in "$@"; do export "$kv"; done
CMD="python tools/exp_fused.py 2048 $LG NAPAFFT_FUSED=1 2"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo plain run failed; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
cat gpurun_out/plain_$TAG.log
# launches per variant round: 3 x (3 warm-up + 2 timed); take the timed natural-order, bit-reversed and inverse ones
ncu --set full --clock-control none --import-source on -k regex:fft_fused -s 3 -c 12 -o /tmp/prof_$TAG -f $CMD > gpurun_out/ncu_$TAG.log 2>&1
for k in 1 6 11; do python tools/ncu_summary.py /tmp/prof_$TAG.ncu-rep $k 60 > gpurun_out/ncu_${TAG}_launch$k.txt 2>&1; done
