#!/bin/bash
mkdir -p gpurun_out
VARIANTS='[{},{"inv_trim":0},{"inv_trim":0,"inv_merge_n":0}]' CTX=bcast timeout 600 python tools/inv_ab.py cfg3 16 75776 > gpurun_out/r2_ab17.log 2>&1
cat gpurun_out/r2_ab17.log
: > gpurun_out/r2_traffic_matrix.log
for cfg in "8 1" "8 2" "6 1" "6 2" "4 1" "4 2"; do
  set -- $cfg
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:flow_tc_inv -s 1 -c 1 --csv --log-file gpurun_out/tmp_traffic.csv python tools/traffic_probe.py 480 500224 $1 $2 >> gpurun_out/r2_traffic_matrix.log 2>&1
  grep -E "dram__bytes|gpu__time" gpurun_out/tmp_traffic.csv | awk -F'","' '{print $13, $15}' >> gpurun_out/r2_traffic_matrix.log
done
cat gpurun_out/r2_traffic_matrix.log
