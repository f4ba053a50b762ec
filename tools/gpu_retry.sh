#!/bin/bash
# usage: gpu_retry.sh <script> <timeout> <outfile>
for i in 1 2 3 4 5 6 7 8 9 10; do
  gpurun --timeout $2 -- "bash $1" > $3 2>&1
  rc=$?
  if grep -q "status=transient" $3; then sleep 150; continue; fi
  break
done
echo "done rc=$rc" >> $3
