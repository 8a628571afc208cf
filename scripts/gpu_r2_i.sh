#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 120 python scripts/gemm_tile_invariance.py > gpurun_out/r2i_tile_invariance.log 2>&1; echo "invariance rc=$?"; cat gpurun_out/r2i_tile_invariance.log
bash scripts/gpu_profile_r2.sh
