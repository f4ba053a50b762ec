#!/bin/bash
# round-2 record: full GPU suite, smoke, default bench, launch list of the same command
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r2_t19.log 2>&1; echo "tests rc=$?" > gpurun_out/r2_rc19.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2_smoke19.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_rc19.log
timeout 1500 python bench.py > gpurun_out/r2_bench19.json 2> gpurun_out/r2_bench19.err; echo "bench rc=$?" >> gpurun_out/r2_rc19.log
timeout 1500 python bench.py --impl reference > gpurun_out/r2_bench19_ref.json 2> gpurun_out/r2_bench19_ref.err; echo "ref rc=$?" >> gpurun_out/r2_rc19.log
timeout 1800 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'flow_|lse_|pack_|fold_|importance|hist|hpd|trunc' -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_launch19.log 2>&1; echo "launchlist rc=$?" >> gpurun_out/r2_rc19.log
tail -n 8 gpurun_out/r2_t19.log; tail -n 12 gpurun_out/r2_smoke19.log; cat gpurun_out/r2_rc19.log; head -c 600 gpurun_out/r2_bench19.json; tail -3 gpurun_out/r2_bench19.err
