#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "variants_headline or context_fold or log_prob_and_sample or golden or reference_outputs" > gpurun_out/r2_t14.log 2>&1; echo "tests rc=$?" > gpurun_out/r2_rc14.log
VARIANTS='[{"inv_kernel":5,"inv_align":0},{"inv_kernel":5},{},{"inv_merge_n":256},{"inv_align":0}]' timeout 600 python tools/inv_ab.py cfg3 16 75776 > gpurun_out/r2_ab14.log 2>&1
for c in cfg4 cfg2 cfg5a; do VARIANTS='[{"inv_kernel":5,"inv_align":0},{"inv_kernel":5},{}]' timeout 300 python tools/inv_ab.py $c 16 37888 >> gpurun_out/r2_ab14.log 2>&1; done
tail -n 15 gpurun_out/r2_t14.log; cat gpurun_out/r2_ab14.log gpurun_out/r2_rc14.log
