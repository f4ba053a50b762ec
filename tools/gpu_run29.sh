#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/inv5_spread.py 3 > gpurun_out/r2_spread29.log 2>&1; echo "rc=$?"
head -60 gpurun_out/r2_spread29.log | cut -c1-260
