#!/bin/bash
mkdir -p gpurun_out
timeout 300 python examples/calibrate_synthetic.py 256 20000 > gpurun_out/r2_example61.log 2>&1; echo "example rc=$?"; tail -6 gpurun_out/r2_example61.log
timeout 600 python -m pytest tests/test_grad.py -q -m gpu -k "training or embedding or train_driver" > gpurun_out/r2_t61.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_t61.log
