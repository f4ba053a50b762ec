#!/bin/bash
# final record of round 2: full GPU suite, smoke, default bench, reference arm, launch list of the bench command
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r2_t62.log 2>&1; echo "tests rc=$?" > gpurun_out/r2_rc62.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2_smoke62.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_rc62.log
timeout 1500 python bench.py > gpurun_out/r2_bench62.json 2> gpurun_out/r2_bench62.err; echo "bench rc=$?" >> gpurun_out/r2_rc62.log
timeout 1500 python bench.py --impl reference > gpurun_out/r2_bench62_ref.json 2> gpurun_out/r2_bench62_ref.err; echo "ref rc=$?" >> gpurun_out/r2_rc62.log

tail -n 4 gpurun_out/r2_t62.log; tail -n 3 gpurun_out/r2_smoke62.log; cat gpurun_out/r2_rc62.log; head -c 300 gpurun_out/r2_bench62.json
