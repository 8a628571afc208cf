#!/bin/bash
set -u
mkdir -p gpurun_out
run() { local name=$1; shift; echo "== $name"; timeout "$@" > "gpurun_out/r2g_$name.log" 2>&1; echo "rc=$? ($(tail -n 1 gpurun_out/r2g_$name.log | cut -c1-300))"; }
run suite 900 python -m pytest tests -q -m gpu -rxXs
run env_step 300 python scripts/env_step_bench.py --batches 1,16,256 --steps 48
run bench 500 python bench.py --no-cpu-baseline
D4_SPACE_V=2 run bench_space_v2 500 python bench.py --no-cpu-baseline --steps 2 --warmup 2
run bench_config3 400 python bench.py --no-cpu-baseline --workload config3 --steps 2 --warmup 2
run ncu_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 1500 --csv --log-file gpurun_out/r2g_launches_b2048.csv python bench.py --horizon 6 --steps 1 --warmup 1 --no-cpu-baseline --no-profile
