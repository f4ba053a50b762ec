#!/bin/bash
mkdir -p gpurun_out
{
python tools/grad_time.py cfg4 4 100000
python tools/grad_time.py cfg2 2 50000
python tools/grad_time.py cfg3 2 50000
python tools/grad_time.py cfg5b 1 20000
} > gpurun_out/r2_gradtime52.log 2>&1
cat gpurun_out/r2_gradtime52.log
python tools/grad_time.py cfg3 2 4096 > /dev/null 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:flow_grad_kernel -c 1 -o gpurun_out/r2_flow_grad_spline python tools/grad_time.py cfg3 2 4096 > gpurun_out/r2_ncu52.log 2>&1; echo "ncu rc=$?"
