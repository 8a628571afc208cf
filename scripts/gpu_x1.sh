#!/bin/bash
set -u
mkdir -p gpurun_out
for P in tf32 tf32x3; do
timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg --clock-control none -c 260 --csv --log-file gpurun_out/launches_${P}.csv \
    python bench.py --batch 2048 --horizon 1 --steps 1 --warmup 0 --no-cpu-baseline --no-profile --precision $P > gpurun_out/ncu_launches_${P}.log 2>&1
echo "launch list $P rc=$?"
done
