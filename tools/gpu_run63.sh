#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_inv.py 32 303104 4 > gpurun_out/r2_prof63_plain.log 2>&1; echo "plain rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flow_tc_inv5 -c 1 -f -o gpurun_out/r2_inv5_final2 python tools/prof_inv.py 32 303104 4 > gpurun_out/r2_ncu63.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/r2_ncu63.log
