#!/bin/bash
# programmatic dependent launch on the GEMM and the three attention kernels of a layer: parity, A/B at 256 and 2048 dreams, env step
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_horizon_parity_gpu.py tests/test_zx_graph_replay_gpu.py -x -q 2>&1 | tail -4
for pdl in 0 1 0 1; do
D4_PDL=$pdl timeout 600 python bench.py --batch 256 --horizon 16 --steps 3 --warmup 3 --no-cpu-baseline --no-weak --no-profile > gpurun_out/r2y_b256_$pdl.json 2> gpurun_out/r2y_b256_$pdl.err; python - <<PY
import json; d=json.loads(open('gpurun_out/r2y_b256_$pdl.json').read().strip().splitlines()[-1]); print('B=256 pdl $pdl', round(d['value'],1), d['ms_per_step'])
PY
done
for pdl in 0 1 0 1; do
D4_PDL=$pdl timeout 600 python bench.py --horizon 16 --steps 3 --warmup 3 --no-cpu-baseline --no-weak --no-profile > gpurun_out/r2y_b2048_$pdl.json 2> gpurun_out/r2y_b2048_$pdl.err; python - <<PY
import json; d=json.loads(open('gpurun_out/r2y_b2048_$pdl.json').read().strip().splitlines()[-1]); print('B=2048 pdl $pdl', round(d['value'],1), d['ms_per_step'])
PY
done
for pdl in 0 1; do D4_PDL=$pdl timeout 300 python scripts/env_step_bench.py --batches 1,16 2>&1 | grep '^{' | cut -c1-220; done
