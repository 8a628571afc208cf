#!/bin/bash
set -u
mkdir -p gpurun_out
run() { local name=$1; shift; echo "== $name"; timeout "$@" > "gpurun_out/r2c_$name.log" 2>&1; echo "rc=$? ($(tail -n 1 gpurun_out/r2c_$name.log | cut -c1-300))"; }
rm -f gpurun_out/gemm_bench.jsonl
D4_EXPERIMENTAL=1 run f16_tests 300 python -m pytest tests/test_zz_gemm_f16_gpu.py -q -m gpu -rxX -x -k test_linear
run gemm_bench 500 python scripts/gemm_bench.py --rows 30720 --modes tf32x3,f16x3,f16x3_noepi,f16x3_nosplit,f16x3_nomma,f16x3_only_ldg,f16x3_old,f16x3_old_noepi,f16x3_old_nosplit,f16x3_old_nomma,f16x3_old_only_tma
