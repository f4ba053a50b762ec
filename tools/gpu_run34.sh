#!/bin/bash
mkdir -p gpurun_out
VARIANTS='[{}]' timeout 600 python tools/inv_ab.py cfg3 16 75776 > gpurun_out/r2_ab34.log 2>&1
for c in cfg4 cfg2 cfg5a; do VARIANTS='[{}]' timeout 300 python tools/inv_ab.py $c 16 37888 >> gpurun_out/r2_ab34.log 2>&1; done
cat gpurun_out/r2_ab34.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "variants_headline or context_fold or log_prob_and_sample or golden or reference_outputs or ragged" > gpurun_out/r2_t34.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_t34.log
