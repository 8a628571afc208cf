#!/bin/bash
# 8-GPU strong-scaling line of the bench contract (2048 dreams sharded 8 x 256, + the weak key), as the driver launches it
set -u
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_n8_bench.log 2> gpurun_out/r2_n8_bench.err
echo "rc=$?"; tail -n 1 gpurun_out/r2_n8_bench.log | cut -c1-900; tail -n 3 gpurun_out/r2_n8_bench.err
