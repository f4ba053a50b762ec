#!/bin/bash
# final record of round 2: full GPU suite, smoke, default bench, reference arm, launch list of the bench command
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r2_t60.log 2>&1; echo "tests rc=$?" > gpurun_out/r2_rc60.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2_smoke60.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_rc60.log
timeout 1500 python bench.py > gpurun_out/r2_bench60.json 2> gpurun_out/r2_bench60.err; echo "bench rc=$?" >> gpurun_out/r2_rc60.log
timeout 1500 python bench.py --impl reference > gpurun_out/r2_bench60_ref.json 2> gpurun_out/r2_bench60_ref.err; echo "ref rc=$?" >> gpurun_out/r2_rc60.log
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'flow_|lse_|pack_|fold_|importance|hist|hpd|trunc|transpose' -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_launch60.log 2>&1; echo "launchlist rc=$?" >> gpurun_out/r2_rc60.log
tail -n 4 gpurun_out/r2_t60.log; tail -n 3 gpurun_out/r2_smoke60.log; cat gpurun_out/r2_rc60.log; head -c 300 gpurun_out/r2_bench60.json
