#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/inv4_timeline.py bcast 3 > gpurun_out/r2_tl_bcast.log 2>&1; echo "tl1 rc=$?" > gpurun_out/r2_rc3.log
timeout 300 python tools/inv4_timeline.py point 3 > gpurun_out/r2_tl_point.log 2>&1; echo "tl2 rc=$?" >> gpurun_out/r2_rc3.log
cat gpurun_out/r2_rc3.log; head -3 gpurun_out/r2_tl_bcast.log
