#!/bin/bash
# ncu evidence for one bench workload (run on the GPU box through gpurun; one GPU).
# usage: tools/profile_workload.sh <c2|c3|c4> <launches per step> [tag]
# Each ncu pass runs only after the same command exited 0 without ncu.
WL=${1:-c3}; NL=${2:-32}; TAG=${3:-r02}
CMD="python bench.py --workload $WL --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
$CMD > gpurun_out/${TAG}_plain_$WL.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain_$WL.log; exit 1; }
tail -1 gpurun_out/${TAG}_plain_$WL.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_$WL.csv $CMD > gpurun_out/${TAG}_ncu_launches_$WL.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches_$WL.csv > gpurun_out/${TAG}_launches_$WL.txt 2>&1
# one full capture of every launch of ONE timed step (skip the warm-up steps' launches)
ncu --set full --clock-control none --import-source on -k "regex:fft_|rfft_post|zip2|split2|bitrev" -s $((3 * NL)) -c $NL -o /tmp/full_$WL -f $CMD > gpurun_out/${TAG}_ncu_full_$WL.log 2>&1
python tools/ncu_traffic.py /tmp/full_$WL.ncu-rep $WL "$CMD" > gpurun_out/${TAG}_traffic_$WL.json 2> gpurun_out/${TAG}_traffic_$WL.err
for k in 1 $((NL / 4 + 1)) $((NL / 2 + 1)) $((3 * NL / 4 + 1)); do python tools/ncu_summary.py /tmp/full_$WL.ncu-rep $k 14 > gpurun_out/${TAG}_ncu_full_${WL}_launch$k.txt 2>&1; done
