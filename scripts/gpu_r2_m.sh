#!/bin/bash
set -u
mkdir -p gpurun_out
run() { local name=$1; shift; echo "== $name"; timeout "$@" > "gpurun_out/r2m_$name.log" 2>&1; echo "rc=$? ($(tail -n 1 gpurun_out/r2m_$name.log | cut -c1-300))"; }
run suite 900 python -m pytest tests -q -m gpu -rxXs
run bench 500 python bench.py --no-cpu-baseline --no-profile --steps 2 --warmup 2
export D4_TRIM_CONE=0
run bench_nocone 500 python bench.py --no-cpu-baseline --no-profile --steps 2 --warmup 2
unset D4_TRIM_CONE
run bench_config3 400 python bench.py --no-cpu-baseline --workload config3 --steps 2 --warmup 2 --no-profile
run bench_config2 300 python bench.py --no-cpu-baseline --workload config2 --steps 4 --warmup 3 --no-profile
run bench_b256 300 python bench.py --no-cpu-baseline --batch 256 --steps 4 --warmup 3 --no-profile
