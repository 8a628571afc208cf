#!/bin/bash
set -u
mkdir -p gpurun_out
run() { local name=$1; shift; echo "== $name"; timeout "$@" > "gpurun_out/r2d_$name.log" 2>&1; echo "rc=$? ($(tail -n 1 gpurun_out/r2d_$name.log | cut -c1-300))"; }
run suite 900 python -m pytest tests -q -m gpu -x -rxXs
D4_GRAPH=1 run env_step_graph 200 python scripts/env_step_bench.py --batches 1,256
run ncu_env 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 600 --csv --log-file gpurun_out/r2d_env_launches.csv python scripts/env_step_bench.py --batches 1 --steps 8 --warmup 4
