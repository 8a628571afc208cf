#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest.log
for E in 0 1; do
echo "== early_mma=$E"; D4_GEMM_EARLY_MMA=$E timeout 600 python bench.py --precision tf32x3 --horizon 8 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_e$E.log 2>&1; echo rc=$?
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_e$E.log').read().strip().splitlines()[-1])
print('frames/s', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'phases', {k:round(v,1) for k,v in d['phase_ms_per_step'].items()}, 'gemm TF', round(d['roofline_gemm']['achieved'],1), 'share', d['kernel_class_share'], 'clk', d['clocks']['sm_mhz'])
PY
done
D4_GEMM_EARLY_MMA=${BEST:-0} timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_tf32x3.csv \
    python bench.py --batch 2048 --horizon 2 --steps 1 --warmup 0 --no-cpu-baseline --no-profile --precision tf32x3 > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
