#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x > gpurun_out/r2_t36.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2_t36.log
VARIANTS='[{}]' timeout 600 python tools/inv_ab.py cfg3 16 75776 > gpurun_out/r2_ab36.log 2>&1
for c in cfg4 cfg2 cfg5a cfg5b cfg5c; do VARIANTS='[{}]' timeout 300 python tools/inv_ab.py $c 16 37888 >> gpurun_out/r2_ab36.log 2>&1; done
cat gpurun_out/r2_ab36.log | cut -c1-150
