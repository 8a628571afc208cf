#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest.log
echo "== bench"; timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_tf32x3.log 2>&1; echo rc=$?; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_tf32x3.log').read().strip().splitlines()[-1])
print('frames/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'phases', {k:round(v,1) for k,v in d['phase_ms_per_step'].items()}, 'gemm TF', round(d['roofline_gemm']['achieved'],1), 'K1', round(d['roofline_attn']['frac'],3), 'share', {k:round(v,3) for k,v in d['kernel_class_share'].items()}, 'clk', d['clocks']['sm_mhz'])
PY
echo "== bench B=256"; timeout 900 python bench.py --batch 256 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_b256.log 2>&1; echo rc=$?; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_b256.log').read().strip().splitlines()[-1])
print('frames/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'gemm TF', round(d['roofline_gemm']['achieved'],1), 'K1', round(d['roofline_attn']['frac'],3), 'share', {k:round(v,3) for k,v in d['kernel_class_share'].items()})
PY
