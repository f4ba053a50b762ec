#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_grad.py -q -m gpu > gpurun_out/r2_t57.log 2>&1; echo "tests rc=$?"
tail -n 5 gpurun_out/r2_t57.log
{
python tools/grad_time.py cfg4 4 100000
python tools/grad_time.py cfg2 2 50000
python tools/grad_time.py cfg3 2 50000
} > gpurun_out/r2_gradtime57.log 2>&1
cat gpurun_out/r2_gradtime57.log
