#!/bin/bash
set -u
mkdir -p gpurun_out
run() { local name=$1; shift; echo "== $name"; timeout "$@" > "gpurun_out/r2l_$name.log" 2>&1; echo "rc=$? ($(tail -n 1 gpurun_out/r2l_$name.log | cut -c1-300))"; }
run trim_test 300 python -m pytest tests/test_horizon_parity_gpu.py -q -m gpu -k trimmed
run bench 500 python bench.py --no-cpu-baseline --no-profile --steps 2 --warmup 2
export D4_TRIM_FINAL=0
run bench_notrim 500 python bench.py --no-cpu-baseline --no-profile --steps 2 --warmup 2
unset D4_TRIM_FINAL
