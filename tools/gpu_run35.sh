#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "n8 rc=$?"
grep '^{' gpurun_out/r2_bench_n8.json | head -c 400; tail -3 gpurun_out/r2_bench_n8.err
