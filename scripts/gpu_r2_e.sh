#!/bin/bash
set -u
mkdir -p gpurun_out
run() { local name=$1; shift; echo "== $name"; timeout "$@" > "gpurun_out/r2e_$name.log" 2>&1; echo "rc=$? ($(tail -n 1 gpurun_out/r2e_$name.log | cut -c1-400))"; }
run suite 900 python -m pytest tests -q -m gpu -rxXs
D4_GRAPH=1 run env_step_graph 200 python scripts/env_step_bench.py --batches 1,16
run env_step 200 python scripts/env_step_bench.py --batches 1,16,256
run bench_b256 300 python bench.py --no-cpu-baseline --precision f16x3 --batch 256 --steps 3 --warmup 2
run ncu_b256 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1200 --csv --log-file gpurun_out/r2e_b256_launches.csv python bench.py --precision f16x3 --batch 256 --horizon 8 --steps 1 --warmup 1 --no-cpu-baseline --no-profile
