#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest linear (tc3)"; timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider -k "linear or midsize" > gpurun_out/pytest_tc3.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_tc3.log
echo "== pytest all"; timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest.log
P=tf32x3
echo "== bench $P"; timeout 900 python bench.py --precision $P --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_${P}.log 2>&1; echo rc=$?; tail -c 2200 gpurun_out/bench_${P}.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_${P}.csv \
    python bench.py --batch 2048 --horizon 2 --steps 1 --warmup 0 --no-cpu-baseline --no-profile --precision $P > gpurun_out/ncu_launches_${P}.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:time_attn_bulk -c 2 -f -o gpurun_out/prof_k1b python scripts/k1_bench.py --ts 1,40 --iters 0 > gpurun_out/ncu_k1b.log 2>&1
echo "k1 ncu rc=$?"
