#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/nsc_tc_probe.py > gpurun_out/r2_nsc_probe.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/r2_nsc_probe.log
