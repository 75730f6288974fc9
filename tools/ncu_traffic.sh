#!/bin/bash
# tools/ncu_traffic.sh : one ncu run per (bench id, log size) pair of bench.py's roofline_hbm_kernels
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum
for cfg in 2:20 3:20 4:23 5:20 1:23 4:27 2:24; do
  w=${cfg%%:*}; l=${cfg##*:}
  timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/traffic_${w}_${l}.csv python tools/ncu_traffic.py run $w $l 2>&1 | tail -1
done
