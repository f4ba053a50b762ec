#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "variants_headline or context_fold or bounded_broadcast or log_prob_and_sample or ragged_and_tiny_batches or golden" > gpurun_out/r2_t10.log 2>&1; echo "tests rc=$?" > gpurun_out/r2_rc10.log
VARIANTS='[{"inv_kernel":4},{"inv_a_tmem":0},{}]' timeout 600 python tools/inv_ab.py cfg3 16 75776 > gpurun_out/r2_ab10.log 2>&1; echo "ab rc=$?" >> gpurun_out/r2_rc10.log
timeout 300 python tools/inv5_timeline.py bcast 3 > gpurun_out/r2_tl10_bcast.log 2>&1; echo "tl rc=$?" >> gpurun_out/r2_rc10.log
tail -n 12 gpurun_out/r2_t10.log; cat gpurun_out/r2_ab10.log gpurun_out/r2_rc10.log; head -2 gpurun_out/r2_tl10_bcast.log
