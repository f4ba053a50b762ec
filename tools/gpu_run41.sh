#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/inv5_timeline.py bcast 3 > gpurun_out/r2_tl41_hack.log 2>&1; echo "rc=$?"
sed -n 1,30p gpurun_out/r2_tl41_hack.log | cut -c1-60
