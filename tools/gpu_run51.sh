#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_grad.py tests/test_gpu_parity.py -q -m gpu -k "grad or training or api_drop_in or permute or train_driver" > gpurun_out/r2_t51.log 2>&1; echo "tests rc=$?"
tail -n 30 gpurun_out/r2_t51.log
