#!/bin/bash
mkdir -p gpurun_out
VARIANTS='[{},{"inv_merge_n":0}]' CTX=bcast timeout 600 python tools/inv_ab.py cfg3 16 75776 > gpurun_out/r2_ab20.log 2>&1
for c in cfg2 cfg5a cfg4; do VARIANTS='[{"inv_merge_n":256},{"inv_merge_n":0}]' timeout 300 python tools/inv_ab.py $c 16 37888 >> gpurun_out/r2_ab20.log 2>&1; done
cat gpurun_out/r2_ab20.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "variants_headline or log_prob_and_sample" > gpurun_out/r2_t20.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_t20.log
timeout 1800 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'flow_|lse_|pack_|fold_|importance|hist|hpd|trunc' -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_launch20.log 2>&1; echo "launchlist rc=$?"
grep -c flow_tc_inv gpurun_out/launches_r2.csv
