#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_stats.py -q -m gpu > gpurun_out/r2_t64.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2_t64.log
