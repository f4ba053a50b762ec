#!/bin/bash
mkdir -p gpurun_out
VARIANTS='[{},{"inv_park":0}]' timeout 600 python tools/inv_ab.py cfg3 16 75776 > gpurun_out/r2_ab30.log 2>&1
for c in cfg4 cfg2 cfg5a; do VARIANTS='[{},{"inv_park":0}]' CTX=bcast timeout 300 python tools/inv_ab.py $c 16 37888 >> gpurun_out/r2_ab30.log 2>&1; done
cat gpurun_out/r2_ab30.log
timeout 300 python tools/inv5_spread.py 3 > gpurun_out/r2_spread30.log 2>&1; head -14 gpurun_out/r2_spread30.log | cut -c1-200
