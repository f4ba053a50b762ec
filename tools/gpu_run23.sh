#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "variants_headline" > gpurun_out/r2_t23.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_t23.log
VARIANTS='[{},{"inv_merge_n":0},{"inv_merge_n":0,"inv_defer":1}]' CTX=bcast timeout 600 python tools/inv_ab.py cfg3 16 75776 > gpurun_out/r2_ab23.log 2>&1
for c in cfg4 cfg2 cfg5a; do VARIANTS='[{},{"inv_merge_n":0,"inv_defer":1}]' CTX=bcast timeout 300 python tools/inv_ab.py $c 16 37888 >> gpurun_out/r2_ab23.log 2>&1; done
cat gpurun_out/r2_ab23.log
timeout 300 python bench.py --steps 1 --warmup 3 --draws 16 --points 131072 --no-cpu-baseline --no-aux > gpurun_out/r2_bench23_small.json 2> gpurun_out/r2_bench23_small.err; echo "small bench rc=$?"; tail -c 700 gpurun_out/r2_bench23_small.json
