#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm|ln_act|sample_actions' -s 1500 -c 700 --csv --log-file gpurun_out/r2ae_launches.csv \
    python bench.py --horizon 4 --steps 1 --warmup 1 --no-cpu-baseline --no-profile --no-weak > gpurun_out/r2ae_launches.log 2>&1; echo "launch list rc=$?"
