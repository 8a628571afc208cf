#!/bin/bash
set -u
mkdir -p gpurun_out
P=tf32x3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc3 -s 2 -c 4 -f -o gpurun_out/prof_gemm3 \
    python bench.py --batch 2048 --horizon 1 --steps 1 --warmup 0 --no-cpu-baseline --no-profile --precision $P > gpurun_out/ncu_gemm3.log 2>&1
echo "gemm ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"l2s_fused|lp_fused|space_attn" -c 3 -f -o gpurun_out/prof_pools \
    python bench.py --batch 2048 --horizon 1 --steps 1 --warmup 0 --no-cpu-baseline --no-profile --precision $P > gpurun_out/ncu_pools.log 2>&1
echo "pools ncu rc=$?"
