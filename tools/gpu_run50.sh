#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "permute_and_batchnorm or python_api_drop_in" > gpurun_out/r2_t50.log 2>&1; echo "tests rc=$?"
tail -n 25 gpurun_out/r2_t50.log
