#!/bin/bash
# One iteration on the GPU box: parity suite, K1 sweep, the full-size bench (tf32x3) and a launch list of one short run.
set -u
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest.log
echo "== k1"; timeout 300 python scripts/k1_bench.py > gpurun_out/k1_bench.log 2>&1; echo rc=$?; head -14 gpurun_out/k1_bench.log
P=${PREC:-tf32x3}
echo "== bench $P"; timeout 900 python bench.py --precision $P --steps ${STEPS:-1} --warmup 1 --no-cpu-baseline ${BENCH_EXTRA:-} > gpurun_out/bench_${P}.log 2>&1; echo rc=$?; tail -c 2500 gpurun_out/bench_${P}.log
if [ "${LAUNCHES:-1}" = "1" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_${P}.csv \
    python bench.py --batch 2048 --horizon 2 --steps 1 --warmup 0 --no-cpu-baseline --no-profile --precision $P > gpurun_out/ncu_launches_${P}.log 2>&1
echo "launch list rc=$?"
fi
