#!/bin/bash
mkdir -p gpurun_out
VARIANTS='[{}]' timeout 600 python tools/inv_ab.py cfg3 16 75776 > gpurun_out/r2_ab32.log 2>&1
for c in cfg4 cfg2; do VARIANTS='[{}]' CTX=bcast timeout 300 python tools/inv_ab.py $c 16 37888 >> gpurun_out/r2_ab32.log 2>&1; done
cat gpurun_out/r2_ab32.log
