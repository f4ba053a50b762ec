#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "variants_headline or context_fold or log_prob_and_sample or golden or reference_outputs or ragged" > gpurun_out/r2_t33.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_t33.log
