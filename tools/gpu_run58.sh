#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2_bench58_n2.json 2> gpurun_out/r2_bench58_n2.err; echo "bench n2 rc=$?"
head -c 300 gpurun_out/r2_bench58_n2.json; tail -3 gpurun_out/r2_bench58_n2.err
