#!/bin/bash
set -u
mkdir -p gpurun_out
run() { local name=$1; shift; echo "== $name"; timeout "$@" > "gpurun_out/r2h_$name.log" 2>&1; echo "rc=$? ($(tail -n 1 gpurun_out/r2h_$name.log | cut -c1-300))"; }
run suite 900 python -m pytest tests -q -m gpu -rxXs
run bench 500 python bench.py --no-cpu-baseline
D4_TRIM_FINAL=0 run bench_notrim 500 python bench.py --no-cpu-baseline --steps 2 --warmup 2 --no-profile
run bench_b256 300 python bench.py --no-cpu-baseline --batch 256 --steps 4 --warmup 3
rm -f gpurun_out/gemm_bench.jsonl
run gemm_same 300 python scripts/gemm_bench.py --rows 30720 --modes f16x3,f16x3_only_tma,f16x3_only_tma_sameA,f16x3_only_tma_sameW,f16x3_only_tma_sameAW --reps 10
