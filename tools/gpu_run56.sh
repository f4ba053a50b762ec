#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r2_t56.log 2>&1; echo "tests rc=$?"
tail -n 6 gpurun_out/r2_t56.log
