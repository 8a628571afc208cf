#!/bin/bash
# Final evidence run: launch list of a short run of the bench command, one --set full capture of K1 (t = 40, alone) and of
# the tcgen05 GEMM launches of one pass.  Outputs in gpurun_out/ (summaries are exported to profiles/ by scripts/ncu_export.sh).
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/final_launches.csv \
    python bench.py --horizon 2 --steps 1 --warmup 0 --no-cpu-baseline --no-profile > gpurun_out/final_launches.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:time_attn_bulk -c 1 -f -o gpurun_out/final_k1 \
    python scripts/k1_bench.py --ts 40 --iters 0 > gpurun_out/final_k1.log 2>&1
echo "k1 capture rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc3 -s 2 -c 6 -f -o gpurun_out/final_gemm \
    python bench.py --horizon 1 --steps 1 --warmup 0 --no-cpu-baseline --no-profile > gpurun_out/final_gemm.log 2>&1
echo "gemm capture rc=$?"
