#!/bin/bash
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/prof_inv.py 2 1000 1 > gpurun_out/r2_memcheck_inv.log 2>&1; echo "memcheck inv rc=$?"
tail -8 gpurun_out/r2_memcheck_inv.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "ragged_and_tiny or golden" > gpurun_out/r2_memcheck_tests.log 2>&1; echo "memcheck tests rc=$?"
tail -8 gpurun_out/r2_memcheck_tests.log
