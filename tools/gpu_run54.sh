#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "coupling" -s > gpurun_out/r2_t54.log 2>&1; echo "tests rc=$?"
tail -n 30 gpurun_out/r2_t54.log
