#!/bin/bash
# DRAM traffic and duration of one pipeline launch (4096 x 2^16, natural order) per chunk size (auto lag/slots)
for kb in ${1:-2048 4096 8192 16384}; do
  export NAPAFFT_PIPE_CHUNK_KB=$kb
  C="python tools/sweep_sizes.py 4096 16"
  plain=$($C | grep "fwd nat")
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:fft_pipeline -s 3 -c 1 --csv $C 2>/dev/null | grep -E "dram__bytes|gpu__time" | awk -F'","' '{printf "%s=%s%s ", $(NF-2), $NF, $(NF-1)}' | tr -d '"'
  echo " | chunk=${kb}KB auto | $plain"
done
