#!/bin/bash
# last check of the round: full GPU suite + smoke on the final build
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r2_t65.log 2>&1; echo "tests rc=$?"
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2_smoke65.log 2>&1; echo "smoke rc=$?"
tail -n 3 gpurun_out/r2_t65.log; tail -n 2 gpurun_out/r2_smoke65.log
