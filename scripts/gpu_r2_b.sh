#!/bin/bash
# Round 2, call B: f16 GEMM v2 (A through registers) - correctness, per-shape timing with ablations, engine parity, bench.
set -u
mkdir -p gpurun_out
run() { local name=$1; shift; echo "== $name"; timeout "$@" > "gpurun_out/r2b_$name.log" 2>&1; echo "rc=$? ($(tail -n 1 gpurun_out/r2b_$name.log | cut -c1-300))"; }
D4_EXPERIMENTAL=1 run f16_tests 300 python -m pytest tests/test_zz_gemm_f16_gpu.py -q -m gpu -rxX -x
run gemm_bench 400 python scripts/gemm_bench.py
run bench_f16x3 300 python bench.py --no-cpu-baseline --precision f16x3
run determinism 300 python scripts/determinism_check.py
