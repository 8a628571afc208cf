#!/bin/bash
# Round 2, call A: first hardware run of the fp16 split GEMM, graph replay, tokenizer throughput.
set -u
mkdir -p gpurun_out
run() { local name=$1; shift; echo "== $name"; timeout "$@" > "gpurun_out/r2_$name.log" 2>&1; echo "rc=$? ($(tail -n 1 gpurun_out/r2_$name.log | cut -c1-300))"; }
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_gpu.txt
D4_EXPERIMENTAL=1 run f16_gemm 300 python -m pytest tests/test_zz_gemm_f16_gpu.py -q -m gpu -k test_linear_f16x3 -rxX
D4_EXPERIMENTAL=1 run f16_engine 300 python -m pytest tests/test_zz_gemm_f16_gpu.py -q -m gpu -k test_f16x3_engine -rxX
run bench_f16x3 400 python bench.py --no-cpu-baseline --precision f16x3
run env_step 200 python scripts/env_step_bench.py --batches 1,16
D4_GRAPH=1 run env_step_graph 200 python scripts/env_step_bench.py --batches 1,16
run tokenizer_bench 200 python scripts/tokenizer_bench.py --batches 4,32
run ncu_f16 400 ncu --set full --clock-control none --import-source on -k regex:gemm_f16x3 -c 2 -o gpurun_out/r2_f16_gemm \
    python bench.py --precision f16x3 --horizon 2 --steps 1 --warmup 0 --no-cpu-baseline --no-profile
ls -la gpurun_out | tail -n 20
