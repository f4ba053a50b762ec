#!/bin/bash
mkdir -p gpurun_out

timeout 120 python tools/hist_time.py > gpurun_out/r2_hist_time.log 2>&1; cat gpurun_out/r2_hist_time.log
