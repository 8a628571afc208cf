set -u
mkdir -p gpurun_out
for P in tf32x3 tf32; do
echo "== bench $P B=256"; timeout 600 python bench.py --precision $P --batch 256 --horizon 64 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_${P}_256.log 2>&1; echo rc=$?; tail -c 3000 gpurun_out/bench_${P}_256.log
echo "== bench $P B=2048"; timeout 900 python bench.py --precision $P --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_${P}_2048.log 2>&1; echo rc=$?; tail -c 3000 gpurun_out/bench_${P}_2048.log
done
PREC=tf32x3 BENCH_B=2048 bash scripts/gpu_profile.sh
