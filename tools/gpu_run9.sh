#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "log_prob_and_sample or golden or sampling or reference_outputs" > gpurun_out/r2_t9.log 2>&1; echo "tests rc=$?" > gpurun_out/r2_rc9.log
ENGS=tcgen05 DIRS=forward timeout 300 python tools/quick_time.py cfg3 1000 10000 > gpurun_out/r2_fwd9.log 2>&1
ENGS=tcgen05 DIRS=forward timeout 300 python tools/quick_time.py cfg5a 1000 10000 >> gpurun_out/r2_fwd9.log 2>&1
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err; echo "bench rc=$?" >> gpurun_out/r2_rc9.log
tail -n 4 gpurun_out/r2_t9.log; cat gpurun_out/r2_fwd9.log gpurun_out/r2_rc9.log; python -c "
import json; d=json.load(open('gpurun_out/r2_bench9.json')); print({k:d[k] for k in ['value','ms_per_step']}, d['roofline']['frac'], d['e2e']['value'], d['aux_sample_direction']['value'], d['clocks'])"
