#!/bin/bash
# N-tile choice of the CTA-pair GEMMs at 256 and 2048 dreams per GPU: default heuristic, pinned 128 / 256, wave model
set -u
mkdir -p gpurun_out
for pin in 0 128 -1 0 -1; do
D4_GEMM_PAIR_BN=$pin timeout 600 python bench.py --batch 256 --horizon 16 --steps 3 --warmup 3 --no-cpu-baseline --no-weak --no-profile > gpurun_out/r2af_b256.json 2> gpurun_out/r2af_b256.err; python - <<PY
import json; d=json.loads(open('gpurun_out/r2af_b256.json').read().strip().splitlines()[-1]); print('B=256 pin $pin', round(d['value'],1), d['ms_per_step'])
PY
done
for pin in 0 -1 0 -1; do
D4_GEMM_PAIR_BN=$pin timeout 600 python bench.py --horizon 16 --steps 3 --warmup 3 --no-cpu-baseline --no-weak --no-profile > gpurun_out/r2af_b2048.json 2> gpurun_out/r2af_b2048.err; python - <<PY
import json; d=json.loads(open('gpurun_out/r2af_b2048.json').read().strip().splitlines()[-1]); print('B=2048 pin $pin', round(d['value'],1), d['ms_per_step'])
PY
done
