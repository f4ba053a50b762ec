#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/inv5_timeline.py bcast 3 > gpurun_out/r2_tl24_unsplit.log 2>&1; echo "rc=$?"
OPTS='{"inv_merge_n":0}' timeout 300 python tools/inv5_timeline.py bcast 3 > gpurun_out/r2_tl24_split.log 2>&1; echo "rc=$?"
head -3 gpurun_out/r2_tl24_unsplit.log; head -3 gpurun_out/r2_tl24_split.log
