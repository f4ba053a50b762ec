#!/bin/bash
mkdir -p gpurun_out
VARIANTS='[{},{"inv_merge_n":0}]' timeout 600 python tools/inv_ab.py cfg3 16 75776 > gpurun_out/r2_ab16.log 2>&1
for c in cfg4 cfg2 cfg5a; do VARIANTS='[{"inv_merge_n":256,"inv_align":0},{}]' timeout 300 python tools/inv_ab.py $c 16 37888 >> gpurun_out/r2_ab16.log 2>&1; done
cat gpurun_out/r2_ab16.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_grad.py -q -m gpu -x -k "variants_headline or twin or bayesian or inverse_grad_matches" > gpurun_out/r2_t16.log 2>&1; echo "tests rc=$?" > gpurun_out/r2_rc16.log
tail -n 5 gpurun_out/r2_t16.log
# full-size DRAM traffic of the headline kernel: two metric passes over ONE launch of the 1000-draw x 1M-point job
timeout 1500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:flow_tc_inv -c 1 --csv --log-file gpurun_out/r2_traffic_full.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-aux --no-parity > gpurun_out/r2_traffic_full.log 2>&1; echo "traffic rc=$?" >> gpurun_out/r2_rc16.log
cat gpurun_out/r2_traffic_full.csv | tail -5; cat gpurun_out/r2_rc16.log
