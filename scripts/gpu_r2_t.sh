#!/bin/bash
# tokenizer transformers in f16x3 (clamp removed): parity, bench, launch list; launch list of the headline bench command
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zy_tokenizer_gpu.py tests/test_zw_tokenizer_fullsize_gpu.py -x -q 2>&1 | tail -5
timeout 300 python scripts/tokenizer_bench.py --batches 32,128 --frames 8 --precision f16x3 > gpurun_out/r2t_tok_f16.log 2>&1; tail -4 gpurun_out/r2t_tok_f16.log | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm|attn|patchify|rstd|rmsnorm|transpose|rows' -s 300 -c 600 --csv --log-file gpurun_out/r2t_tok_launches_f16x3.csv python scripts/tokenizer_bench.py --batches 128 --frames 4 --repeat 2 --precision f16x3 > gpurun_out/r2t_tok_ncu_f16x3.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 1500 --csv --log-file gpurun_out/final_launches.csv \
    python bench.py --horizon 6 --steps 1 --warmup 1 --no-cpu-baseline --no-profile --no-weak > gpurun_out/final_launches.log 2>&1; echo "launch list rc=$?"
