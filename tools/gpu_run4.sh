#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "variants_headline or context_fold" > gpurun_out/r2_t11.log 2>&1; echo "tests rc=$?" > gpurun_out/r2_rc11.log
VARIANTS='[{}]' timeout 600 python tools/inv_ab.py cfg3 16 75776 > gpurun_out/r2_ab11.log 2>&1
for c in cfg4 cfg2 cfg5a cfg5b cfg5c; do VARIANTS='[{"inv_kernel":3},{}]' timeout 300 python tools/inv_ab.py $c 16 37888 >> gpurun_out/r2_ab11.log 2>&1; done
timeout 600 python tools/parity_stats.py > gpurun_out/r2_parity_stats.log 2>&1
tail -n 3 gpurun_out/r2_t11.log; cat gpurun_out/r2_parity_stats.log; cat gpurun_out/r2_ab11.log gpurun_out/r2_rc11.log
