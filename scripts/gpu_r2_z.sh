#!/bin/bash
# small-M behaviour of gemm_f16x3: fixed cost per launch (rows 512 .. 7680), ablations at 3840 rows
set -u
mkdir -p gpurun_out
timeout 600 python scripts/gemm_bench.py --rows 512,1024,2048,3840,7680 --reps 15 --modes f16x3,f16x3_noepi,f16x3_only_tma > gpurun_out/r2z_gemm.log 2>&1; grep -c . gpurun_out/r2z_gemm.log
