#!/bin/bash
# gemm_f16x3 direct (register -> global) epilogue: unit tests, stand-alone A/B vs the staged one, in situ A/B
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_gemm_f16_gpu.py tests/test_gpu_parity.py tests/test_zy_tokenizer_gpu.py -x -q 2>&1 | tail -4
timeout 600 python scripts/gemm_bench.py --rows 30720,8192 --reps 15 --modes f16x3,f16x3_staged,f16x3_nostore,f16x3_noepi > gpurun_out/r2x_gemm.log 2>&1; grep -c . gpurun_out/r2x_gemm.log
for dbg in 32 0 32 0; do
D4_GEMM_F16_DBG=$dbg timeout 600 python bench.py --horizon 16 --steps 3 --warmup 3 --no-cpu-baseline --no-weak --no-profile > gpurun_out/r2x_bench_$dbg.json 2> gpurun_out/r2x_bench_$dbg.err; python - <<PY
import json; d=json.loads(open('gpurun_out/r2x_bench_$dbg.json').read().strip().splitlines()[-1]); print('dbg $dbg', round(d['value'],1), d['ms_per_step'])
PY
done
