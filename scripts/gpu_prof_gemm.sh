#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest all (fused pools on)"; timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest.log
P=tf32x3
# ff-in GEMM (GLU) = 8th gemm_tc3 launch of the first pass, qkvgm = 4th, attn-out (residual) = 5th, ff-out = 9th
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc3 -s 3 -c 6 -f -o gpurun_out/prof_gemm3 \
    python bench.py --batch 2048 --horizon 1 --steps 1 --warmup 0 --no-cpu-baseline --no-profile --precision $P > gpurun_out/ncu_gemm3.log 2>&1
echo "gemm ncu rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_${P}.csv \
    python bench.py --batch 2048 --horizon 1 --steps 1 --warmup 0 --no-cpu-baseline --no-profile --precision $P > gpurun_out/ncu_launches_${P}.log 2>&1
echo "launch list rc=$?"
