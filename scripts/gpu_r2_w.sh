#!/bin/bash
# gemm_f16x3 epilogue: what its TMA stores cost (stand-alone), early residual loads A/B in situ
set -u
mkdir -p gpurun_out
timeout 600 python scripts/gemm_bench.py --rows 30720 --reps 15 --modes f16x3,f16x3_nostore,f16x3_noepi > gpurun_out/r2w_gemm.log 2>&1; grep -c . gpurun_out/r2w_gemm.log
for dbg in 64 0 64 0; do
D4_GEMM_F16_DBG=$dbg timeout 600 python bench.py --horizon 16 --steps 3 --warmup 3 --no-cpu-baseline --no-weak --no-profile > gpurun_out/r2w_bench_$dbg.json 2> gpurun_out/r2w_bench_$dbg.err; python - <<PY
import json; d=json.loads(open('gpurun_out/r2w_bench_$dbg.json').read().strip().splitlines()[-1]); print('dbg $dbg', round(d['value'],1), d['ms_per_step'])
PY
done
