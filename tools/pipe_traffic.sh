#!/bin/bash
# DRAM traffic, L2 hit rate and duration of one group-fused launch (4096 x 2^16 by default) per group size
LG=${LG:-16}
export NAPAFFT_PIPELINE=1
for gs in ${@:-8 16}; do
  export NAPAFFT_PIPE_GSIZE=$gs
  C="python tools/sweep_sizes.py 4096 $LG"
  plain=$($C | grep "fwd nat")
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio --clock-control none -k regex:fft_pipeline -s 3 -c 3 --csv $C 2>/dev/null | grep -E "fft_pipeline" | awk -F'","' '{printf "%s=%s%s ", $(NF-2), $NF, $(NF-1)}' | tr -d '"'
  echo " | gsize=$gs | $plain"
done
