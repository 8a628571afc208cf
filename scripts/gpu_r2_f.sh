#!/bin/bash
set -u
mkdir -p gpurun_out
run() { local name=$1; shift; echo "== $name"; timeout "$@" > "gpurun_out/r2f_$name.log" 2>&1; echo "rc=$? ($(tail -n 1 gpurun_out/r2f_$name.log | cut -c1-400))"; }
run suite 900 python -m pytest tests -q -m gpu -rxXs
run env_step 200 python scripts/env_step_bench.py --batches 1,16,256 --steps 48
D4_GRAPH=0 run env_step_nograph 200 python scripts/env_step_bench.py --batches 1,16
run bench_b256 300 python bench.py --no-cpu-baseline --batch 256 --steps 4 --warmup 3
D4_GRAPH=0 run bench_b256_nograph 300 python bench.py --no-cpu-baseline --batch 256 --steps 4 --warmup 3 --no-profile
run bench_config1 200 python bench.py --no-cpu-baseline --workload config1 --steps 10 --warmup 3
run bench_config2 300 python bench.py --no-cpu-baseline --workload config2 --steps 5 --warmup 3
run bench_config3 400 python bench.py --no-cpu-baseline --workload config3 --steps 3 --warmup 3
