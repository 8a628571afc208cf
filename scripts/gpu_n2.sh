#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest.log
echo "== N=1"; timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_n1.log 2>&1; echo rc=$?; tail -c 1800 gpurun_out/bench_n1.log | head -c 700
echo "== N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_n2.log 2>&1; echo rc=$?; tail -c 2500 gpurun_out/bench_n2.log
