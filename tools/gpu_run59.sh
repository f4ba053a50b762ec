#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/nsc_grad_probe.py > gpurun_out/r2_nsc_grad_probe.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r2_nsc_grad_probe.log
