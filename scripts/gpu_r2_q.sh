#!/bin/bash
# frame_attn_mma.cu: parity + A/B on the tokenizer bench, then the config-5 sweep
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zy_tokenizer_gpu.py -x -q 2>&1 | tail -5
D4_FRAME_MMA=0 timeout 300 python scripts/tokenizer_bench.py --batches 128 --frames 8 > gpurun_out/r2q_tok_fma.log 2>&1; tail -2 gpurun_out/r2q_tok_fma.log | cut -c1-300
timeout 300 python scripts/tokenizer_bench.py --batches 32,128 --frames 8 > gpurun_out/r2q_tok_mma.log 2>&1; tail -4 gpurun_out/r2q_tok_mma.log | cut -c1-300
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 800 --csv --log-file gpurun_out/r2q_tok_launches.csv python scripts/tokenizer_bench.py --batches 128 --frames 4 --repeat 2 > gpurun_out/r2q_tok_ncu.log 2>&1; echo "ncu rc=$?"
timeout 1200 python scripts/sweep.py gpurun_out/r2_sweep_config5.jsonl 2>&1 | tail -10
