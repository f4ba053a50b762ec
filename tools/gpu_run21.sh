#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "n2 rc=$?"
tail -c 1800 gpurun_out/r2_bench_n2.json; tail -5 gpurun_out/r2_bench_n2.err
