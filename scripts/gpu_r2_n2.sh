#!/bin/bash
# 2-GPU validation of the bench contract (strong scaling: the 2048-dream batch sharded over the ranks, + the weak key)
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_n2_bench.log 2> gpurun_out/r2_n2_bench.err
echo "rc=$?"; tail -n 1 gpurun_out/r2_n2_bench.log | cut -c1-600; tail -n 5 gpurun_out/r2_n2_bench.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --cpu-sample 4x4 > gpurun_out/r2_n2_ref.log 2>&1
echo "ref rc=$?"; grep '^{' gpurun_out/r2_n2_ref.log | cut -c1-300
