#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:l2s_fused -s 4 -c 1 -o gpurun_out/r2al_l2s -f \
    python bench.py --horizon 3 --steps 1 --warmup 0 --no-cpu-baseline --no-profile --no-weak > gpurun_out/r2al.log 2>&1; echo "rc=$?"
