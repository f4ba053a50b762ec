#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r2_t39.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_t39.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2_smoke39.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke39.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-aux > gpurun_out/r2_bench39.json 2> gpurun_out/r2_bench39.err; echo "bench rc=$?"; head -c 330 gpurun_out/r2_bench39.json
