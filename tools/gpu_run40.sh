#!/bin/bash
mkdir -p gpurun_out
VARIANTS='[{},{"inv_park":2},{"inv_park":1}]' CTX=bcast timeout 600 python tools/inv_ab.py cfg3 16 75776 > gpurun_out/r2_ab40.log 2>&1
VARIANTS='[{},{"inv_park":2}]' CTX=bcast timeout 600 python tools/inv_ab.py cfg2 16 37888 >> gpurun_out/r2_ab40.log 2>&1
cat gpurun_out/r2_ab40.log
