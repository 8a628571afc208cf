#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"space_attn_mma|l2s_fused|lp_fused|pool_attn" -c 5 -f -o gpurun_out/prof_small \
    python bench.py --horizon 1 --steps 1 --warmup 0 --no-cpu-baseline --no-profile > gpurun_out/ncu_small.log 2>&1
echo "rc=$?"
