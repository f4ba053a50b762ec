#!/bin/bash
# round-2 record (final build of the session): full GPU suite, smoke, default bench, reference arm
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r2_t53.log 2>&1; echo "tests rc=$?" > gpurun_out/r2_rc53.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2_smoke53.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_rc53.log
timeout 1500 python bench.py > gpurun_out/r2_bench53.json 2> gpurun_out/r2_bench53.err; echo "bench rc=$?" >> gpurun_out/r2_rc53.log
timeout 1500 python bench.py --impl reference > gpurun_out/r2_bench53_ref.json 2> gpurun_out/r2_bench53_ref.err; echo "ref rc=$?" >> gpurun_out/r2_rc53.log
tail -n 5 gpurun_out/r2_t53.log; tail -n 6 gpurun_out/r2_smoke53.log; cat gpurun_out/r2_rc53.log; head -c 400 gpurun_out/r2_bench53.json; tail -3 gpurun_out/r2_bench53.err
