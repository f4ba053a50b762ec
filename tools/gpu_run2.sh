#!/bin/bash
mkdir -p gpurun_out
VARIANTS='[{"inv_kernel":3},{},{"inv_wait_hint":10000000},{"inv_wait_hint":2000}]' timeout 600 python tools/inv_ab.py cfg3 16 75776 > gpurun_out/r2_ab2.log 2>&1; echo "ab rc=$?" >> gpurun_out/r2_rc2.log
cat gpurun_out/r2_ab2.log
timeout 300 python tools/prof_inv.py > gpurun_out/prof_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:flow_tc_inv4 -s 1 -c 1 -f -o gpurun_out/inv4_a python tools/prof_inv.py > gpurun_out/ncu_inv4_a.log 2>&1
echo "ncu rc=$?" >> gpurun_out/r2_rc2.log; cat gpurun_out/r2_rc2.log
