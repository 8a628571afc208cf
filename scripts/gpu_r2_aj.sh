#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_horizon_parity_gpu.py -x -q -k "persistent or trimmed" 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'lp_fused' -s 20 -c 20 --csv --log-file gpurun_out/r2aj_launches.csv \
    python bench.py --horizon 6 --steps 1 --warmup 1 --no-cpu-baseline --no-profile --no-weak > gpurun_out/r2aj_launches.log 2>&1; echo "launch list rc=$?"
