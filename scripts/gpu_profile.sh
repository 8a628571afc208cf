#!/bin/bash
# ncu evidence for the two kernels the roofline names.  Outputs in gpurun_out/ (copy summaries to profiles/).
set -u
mkdir -p gpurun_out
P=${PREC:-tf32x3}; B=${BENCH_B:-2048}
# (1) launch list with device time of every kernel of a short run (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_${P}.csv \
    python bench.py --batch $B --horizon 3 --steps 1 --warmup 0 --no-cpu-baseline --no-profile --precision $P > gpurun_out/ncu_launches_${P}.log 2>&1
echo "launch list rc=$?"
# (2) K1 at context t = 40 (frame 41: 40 frames * 5 passes * 2 time layers = 400 earlier K1 launches)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:time_attn -s 400 -c 2 -f -o gpurun_out/prof_k1_${P} \
    python bench.py --batch $B --horizon 42 --steps 1 --warmup 0 --no-cpu-baseline --no-profile --precision $P > gpurun_out/ncu_k1_${P}.log 2>&1
echo "k1 capture rc=$?"
# (3) the tcgen05 GEMM: a few launches from the middle of the first frame
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 60 -c 4 -f -o gpurun_out/prof_gemm_${P} \
    python bench.py --batch $B --horizon 2 --steps 1 --warmup 0 --no-cpu-baseline --no-profile --precision $P > gpurun_out/ncu_gemm_${P}.log 2>&1
echo "gemm capture rc=$?"
ls -la gpurun_out/*.ncu-rep
