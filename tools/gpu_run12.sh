#!/bin/bash
# round-2 baseline record: full GPU suite, default bench, launch list (no full ncu yet)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2_t12.log 2>&1; echo "tests rc=$?" > gpurun_out/r2_rc12.log
timeout 1200 python bench.py > gpurun_out/r2_bench12.json 2> gpurun_out/r2_bench12.err; echo "bench rc=$?" >> gpurun_out/r2_rc12.log
tail -n 6 gpurun_out/r2_t12.log; cat gpurun_out/r2_rc12.log; tail -c 3000 gpurun_out/r2_bench12.json; tail -n 5 gpurun_out/r2_bench12.err
