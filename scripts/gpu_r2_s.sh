#!/bin/bash
# final-state checks: whole GPU suite, ncu captures, the driver's round-end commands, tokenizer launch lists
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
bash scripts/gpu_profile_r2.sh
bash scripts/gpu_round_end.sh
for p in tf32x3 f16x3; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm|attn|patchify|rstd|rmsnorm|transpose|rows' -s 300 -c 600 --csv --log-file gpurun_out/r2s_tok_launches_$p.csv python scripts/tokenizer_bench.py --batches 128 --frames 4 --repeat 2 --precision $p > gpurun_out/r2s_tok_ncu_$p.log 2>&1; echo "ncu $p rc=$?"
done
