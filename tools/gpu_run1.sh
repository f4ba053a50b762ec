#!/bin/bash
# round-2 GPU iteration script: parity of the v4 inverse kernel, then A/B timing (each step under its own timeout)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_smi.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "variants_headline" > gpurun_out/r2_t_variants.log 2>&1; echo "variants rc=$?" >> gpurun_out/r2_rc.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "context_fold or bounded_broadcast" > gpurun_out/r2_t_fold.log 2>&1; echo "fold rc=$?" >> gpurun_out/r2_rc.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "log_prob_and_sample or golden or reference_outputs or ragged" > gpurun_out/r2_t_main.log 2>&1; echo "main rc=$?" >> gpurun_out/r2_rc.log
timeout 600 python tools/inv_ab.py cfg3 16 75776 > gpurun_out/r2_ab1.log 2>&1; echo "ab rc=$?" >> gpurun_out/r2_rc.log
tail -5 gpurun_out/r2_t_variants.log gpurun_out/r2_t_fold.log gpurun_out/r2_t_main.log; cat gpurun_out/r2_ab1.log gpurun_out/r2_rc.log
