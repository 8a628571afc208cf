#!/bin/bash
# final tree: whole GPU suite + the driver's round-end commands + launch list
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
bash scripts/gpu_round_end.sh
