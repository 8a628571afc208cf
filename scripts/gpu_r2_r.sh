#!/bin/bash
# cheaper tf32 split / softclamp in the mma.sync attention kernels: parity, tokenizer A/B (resident CTAs), headline bench
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_zy_tokenizer_gpu.py tests/test_gpu_parity.py tests/test_horizon_parity_gpu.py -x -q 2>&1 | tail -5
D4_FRAME_MINB=1 timeout 300 python scripts/tokenizer_bench.py --batches 128 --frames 8 > gpurun_out/r2r_tok_minb1.log 2>&1; tail -2 gpurun_out/r2r_tok_minb1.log | cut -c1-200
D4_FRAME_MINB=2 timeout 300 python scripts/tokenizer_bench.py --batches 128 --frames 8 > gpurun_out/r2r_tok_minb2.log 2>&1; tail -2 gpurun_out/r2r_tok_minb2.log | cut -c1-200
timeout 300 python scripts/tokenizer_bench.py --batches 128 --frames 8 --precision f16x3 > gpurun_out/r2r_tok_f16.log 2>&1; tail -2 gpurun_out/r2r_tok_f16.log | cut -c1-200
D4_FRAME_MINB=2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 800 --csv --log-file gpurun_out/r2r_tok_launches.csv python scripts/tokenizer_bench.py --batches 128 --frames 4 --repeat 2 --precision f16x3 > gpurun_out/r2r_tok_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 python bench.py > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; cat gpurun_out/r2r_bench.json | cut -c1-600
