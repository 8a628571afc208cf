#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest BK=16"; D4_GEMM_BK=16 timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider -k "linear or midsize or learn_tf32x3" > gpurun_out/pytest_bk16.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_bk16.log
for CFG in "32 0" "16 0" "16 1" "32 1"; do
set -- $CFG
D4_GEMM_BK=$1 D4_GEMM_L2PF=$2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file gpurun_out/launches_bk$1_pf$2.csv \
    python bench.py --batch 2048 --horizon 1 --steps 1 --warmup 0 --no-cpu-baseline --no-profile --precision tf32x3 > gpurun_out/ncu_l.log 2>&1
echo "launch list bk=$1 pf=$2 rc=$?"
done
