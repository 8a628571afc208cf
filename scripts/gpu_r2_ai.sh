#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/sanitize_new_kernels.py 2>&1 | tail -2
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitize_new_kernels.py > gpurun_out/r2ai_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2ai_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python scripts/sanitize_new_kernels.py > gpurun_out/r2ai_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2ai_racecheck.log
