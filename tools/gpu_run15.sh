#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "variants_headline or context_fold or log_prob_and_sample or golden or reference_outputs" > gpurun_out/r2_t15.log 2>&1; echo "tests rc=$?" > gpurun_out/r2_rc15.log
VARIANTS='[{},{"inv_merge_n":0},{"inv_merge_n":96},{"inv_align":1},{"inv_kernel":6,"inv_align":1}]' timeout 600 python tools/inv_ab.py cfg3 16 75776 > gpurun_out/r2_ab15.log 2>&1
for c in cfg4 cfg2 cfg5a; do VARIANTS='[{},{"inv_merge_n":0},{"inv_align":1},{"inv_align":1,"inv_merge_n":0}]' timeout 300 python tools/inv_ab.py $c 16 37888 >> gpurun_out/r2_ab15.log 2>&1; done
tail -n 5 gpurun_out/r2_t15.log; cat gpurun_out/r2_ab15.log gpurun_out/r2_rc15.log
# ncu: the v5 kernel at a size whose packed weights (32 draws x 8.35 MB = 267 MB) exceed L2, 4 draw groups, 16 tiles per CTA
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flow_tc_inv5 -c 1 -f -o gpurun_out/r2_inv5_full python tools/prof_inv.py 32 303104 4 > gpurun_out/r2_ncu15.log 2>&1; echo "ncu rc=$?" >> gpurun_out/r2_rc15.log
tail -3 gpurun_out/r2_ncu15.log
