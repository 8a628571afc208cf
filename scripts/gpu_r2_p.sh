#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/tokenizer_bench.py --batches 32,128 --frames 8 > gpurun_out/r2p_tok_bench.log 2>&1; tail -4 gpurun_out/r2p_tok_bench.log | cut -c1-250
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 800 --csv --log-file gpurun_out/r2p_tok_launches.csv python scripts/tokenizer_bench.py --batches 128 --frames 4 --repeat 2 > gpurun_out/r2p_tok_ncu.log 2>&1; echo "ncu rc=$?"
