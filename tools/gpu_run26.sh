#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x > gpurun_out/r2_t26.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_t26.log
VARIANTS='[{},{"inv_defer":0}]' timeout 600 python tools/inv_ab.py cfg3 16 75776 > gpurun_out/r2_ab26.log 2>&1
for c in cfg4 cfg2 cfg5a; do VARIANTS='[{},{"inv_defer":0}]' timeout 300 python tools/inv_ab.py $c 16 37888 >> gpurun_out/r2_ab26.log 2>&1; done
cat gpurun_out/r2_ab26.log
