#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "variants_headline" > gpurun_out/r2_t25.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_t25.log
VARIANTS='[{},{"inv_merge_n":0},{"inv_merge_n":0,"inv_defer":2}]' CTX=bcast timeout 600 python tools/inv_ab.py cfg3 16 75776 > gpurun_out/r2_ab25.log 2>&1
for c in cfg4 cfg2 cfg5a; do VARIANTS='[{},{"inv_merge_n":0,"inv_defer":2}]' CTX=bcast timeout 300 python tools/inv_ab.py $c 16 37888 >> gpurun_out/r2_ab25.log 2>&1; done
cat gpurun_out/r2_ab25.log
OPTS='{"inv_merge_n":0,"inv_defer":2}' timeout 300 python tools/inv5_timeline.py bcast 3 > gpurun_out/r2_tl25_split_dbl.log 2>&1; head -2 gpurun_out/r2_tl25_split_dbl.log
