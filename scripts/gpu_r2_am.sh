#!/bin/bash
# space attention with sector-friendly fragment loads: parity, launch list, bench
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_horizon_parity_gpu.py -x -q 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'space_attn' -s 20 -c 30 --csv --log-file gpurun_out/r2am_launches.csv \
    python bench.py --horizon 6 --steps 1 --warmup 1 --no-cpu-baseline --no-profile --no-weak > gpurun_out/r2am_launches.log 2>&1; echo "launch list rc=$?"
timeout 600 python bench.py --horizon 16 --steps 3 --warmup 3 --no-cpu-baseline --no-weak --no-profile > gpurun_out/r2am_bench.json 2> gpurun_out/r2am_bench.err; python - <<PY
import json; d=json.loads(open('gpurun_out/r2am_bench.json').read().strip().splitlines()[-1]); print('bench h16', round(d['value'],1), d['ms_per_step'])
PY
