#!/bin/bash
# First gpurun call of round 2: everything drafted in round 1 after its GPU budget ran out gets its first hardware run, cheapest
# and safest first, each step under its own timeout and logged to gpurun_out/ so that a fault in one step costs only that step.
#     gpurun --timeout 1500 -- 'bash scripts/gpu_round2_first.sh'
set -u
mkdir -p gpurun_out
run() { local name=$1; shift; echo "== $name"; timeout "$@" > "gpurun_out/r2_$name.log" 2>&1; echo "rc=$? ($(tail -n 1 gpurun_out/r2_$name.log | cut -c1-160))"; }

# 1. the established suite (includes the non-strict xfail files: full-size properties, graph replay, tokenizer)
run suite 900 python -m pytest tests -q -m gpu -x -rxX
# 2. the fp16 3-term split GEMM stand-alone, then the engine's f16x3 mode against the oracle (may trap: own processes)
D4_EXPERIMENTAL=1 run f16_gemm 300 python -m pytest tests/test_zz_gemm_f16_gpu.py -q -m gpu -k test_linear_f16x3 -rxX
D4_EXPERIMENTAL=1 run f16_engine 300 python -m pytest tests/test_zz_gemm_f16_gpu.py -q -m gpu -k test_f16x3_engine -rxX
# 3. what they buy: the headline bench in both precisions, env-step latency with and without graph replay, tokenizer throughput
run bench_tf32x3 600 python bench.py --no-cpu-baseline
run bench_f16x3 600 python bench.py --no-cpu-baseline --precision f16x3
run env_step 300 python scripts/env_step_bench.py --batches 1,16
D4_GRAPH=1 run env_step_graph 300 python scripts/env_step_bench.py --batches 1,16
run tokenizer_bench 300 python scripts/tokenizer_bench.py --batches 4,32
# 4. one ncu capture of the fp16 GEMM (feed-forward-in shape) for the shared-memory-bandwidth question DESIGN.md section 9 raises
run ncu_f16 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f16x3 -c 2 -o gpurun_out/r2_f16_gemm \
    python bench.py --precision f16x3 --horizon 2 --steps 1 --warmup 0 --no-cpu-baseline --no-profile
ls -la gpurun_out | tail -n 20
