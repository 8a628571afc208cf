#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest.log
echo "== bench"; timeout 900 python bench.py --steps 2 --warmup 2 --cpu-sample 8x8 > gpurun_out/bench_tf32x3.log 2>&1; echo rc=$?; tail -c 1500 gpurun_out/bench_tf32x3.log
echo "== sweep"; timeout 900 python scripts/sweep.py gpurun_out/sweep.jsonl 2>&1 | cut -c1-330
