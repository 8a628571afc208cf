#!/bin/bash
# residual epilogue of gemm_f16x3: L2 prefetch a tile ahead + early chunk loads; unit tests, stand-alone A/B, headline
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_gemm_f16_gpu.py tests/test_gpu_parity.py -x -q 2>&1 | tail -4
timeout 600 python scripts/gemm_bench.py --rows 30720,8192 --reps 15 --modes f16x3,f16x3_res_nol2,f16x3_res_late,f16x3_res_r1,f16x3_noepi > gpurun_out/r2v_gemm.log 2>&1; grep -c . gpurun_out/r2v_gemm.log
timeout 600 python bench.py --horizon 16 --steps 3 --warmup 3 --no-cpu-baseline --no-weak > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; python - <<PY
import json; d=json.loads(open('gpurun_out/r2v_bench.json').read().strip().splitlines()[-1]); print('new', round(d['value'],1), d['ms_per_step'], d['kernel_class_share'], d['roofline']['achieved'])
PY
