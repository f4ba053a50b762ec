#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_grad.py tests/test_gpu_parity.py -q -m gpu -k "grad or batchnorm or training or embedding or train_driver" > gpurun_out/r2_t66.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2_t66.log
