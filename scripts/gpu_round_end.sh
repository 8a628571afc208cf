#!/bin/bash
# What the driver runs at round end, plus the launch list of the same bench command.
set -u
mkdir -p gpurun_out
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo rc=$?; tail -c 600 gpurun_out/final_bench_reference.json
echo "== native arm (defaults)"; timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo rc=$?; tail -c 400 gpurun_out/final_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/final_launches.csv \
    python bench.py --horizon 2 --steps 1 --warmup 0 --no-cpu-baseline --no-profile > gpurun_out/final_launches.log 2>&1
echo "launch list rc=$?"
