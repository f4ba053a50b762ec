#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "variants_headline or context_fold" > gpurun_out/r2_t13.log 2>&1; echo "tests rc=$?" > gpurun_out/r2_rc13.log
VARIANTS='[{"inv_kernel":5},{},{"inv_merge_n":256},{"inv_a_tmem":0}]' timeout 600 python tools/inv_ab.py cfg3 16 75776 > gpurun_out/r2_ab13.log 2>&1
for c in cfg4 cfg2 cfg5a; do VARIANTS='[{"inv_kernel":5},{}]' timeout 300 python tools/inv_ab.py $c 16 37888 >> gpurun_out/r2_ab13.log 2>&1; done
tail -n 5 gpurun_out/r2_t13.log; cat gpurun_out/r2_ab13.log gpurun_out/r2_rc13.log
