#!/bin/bash
# What the driver runs at round end, plus the launch list of the same bench command.
set -u
mkdir -p gpurun_out
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo rc=$?; tail -c 600 gpurun_out/final_bench_reference.json
echo "== native arm (defaults)"; timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo rc=$?; tail -c 400 gpurun_out/final_bench.json
# -s 4000: past the one-time weight packing (torch kernels) and the warm-up step, into the timed step's frames
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 1500 --csv --log-file gpurun_out/final_launches.csv \
    python bench.py --horizon 6 --steps 1 --warmup 1 --no-cpu-baseline --no-profile --no-weak > gpurun_out/final_launches.log 2>&1
echo "launch list rc=$?"
