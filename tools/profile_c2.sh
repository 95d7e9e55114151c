#!/bin/bash
# ncu evidence for the headline workload (run on the GPU box through gpurun; one GPU).
# Each ncu pass runs only after the same command exited 0 without ncu.
set -x
WL=${1:-c2}
CMD="python bench.py --workload $WL --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_$WL.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$WL.csv $CMD > gpurun_out/ncu_launches_$WL.log 2>&1
# one full capture of every distinct launch of a step (skip the warm-up steps' launches)
NL=${2:-4}
ncu --set full --clock-control none --import-source on -k "regex:fft_pass_kernel|rfft_post|zip2|split2" -s $((3 * NL)) -c $NL -o /tmp/full_$WL -f $CMD > gpurun_out/ncu_full_$WL.log 2>&1
for k in $(seq 1 $NL); do python tools/ncu_summary.py /tmp/full_$WL.ncu-rep $k 12 > gpurun_out/ncu_full_${WL}_launch$k.txt 2>&1; done
