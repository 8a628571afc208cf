#!/bin/bash
# Round-2 ncu captures (one B200, config 4, 2048 dreams, f16x3): `--set full` of one layer's worth of gemm_f16x3 launches, of K1 at
# context 40, of the space / pool attention kernels.  Read here with scripts/ncu_summary_r2.py -> profiles/ncu_summary.json.
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
BENCH="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-profile --no-weak"
# f16 GEMM launches of a pass: [l2s out, value residual] + 7 per layer (qkv, attn out, ff in, ff out, pool q, pool kv, pool out): layer 3 = 23..29
timeout 600 $NCU -k regex:gemm_f16x3 -s 23 -c 7 -o gpurun_out/r2_prof_gemm $BENCH --horizon 2 > gpurun_out/r2_prof_gemm.log 2>&1; echo "gemm rc=$?"
# K1: 2 time layers x 5 passes per frame; frame 40 pass 0 = launch 400
timeout 900 $NCU -k regex:time_attn_bulk -s 400 -c 1 -o gpurun_out/r2_prof_k1 $BENCH --horizon 42 > gpurun_out/r2_prof_k1.log 2>&1; echo "k1 rc=$?"
timeout 600 $NCU -k regex:space_attn_reg -s 6 -c 1 -o gpurun_out/r2_prof_space $BENCH --horizon 2 > gpurun_out/r2_prof_space.log 2>&1; echo "space rc=$?"
timeout 600 $NCU -k regex:pool_attn -s 12 -c 1 -o gpurun_out/r2_prof_pool $BENCH --horizon 2 > gpurun_out/r2_prof_pool.log 2>&1; echo "pool rc=$?"
ls -la gpurun_out/r2_prof_*
