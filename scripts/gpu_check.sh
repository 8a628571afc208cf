#!/bin/bash
# One GPU-box visit: smoke, the -m gpu parity suite, short bench runs.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -5 gpurun_out/smoke.log
echo "== pytest" ; timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest.log 2>&1 ; echo "pytest rc=$?" ; tail -40 gpurun_out/pytest.log
for v in 0 1; do
  echo "== bench variant $v" ; timeout 900 python bench.py --batch ${BENCH_B:-64} --horizon ${BENCH_H:-16} --steps 2 --warmup 1 --variant $v --cpu-sample 4x4 ${BENCH_EXTRA:-} > gpurun_out/bench_v$v.log 2>&1 ; echo "bench rc=$?" ; tail -3 gpurun_out/bench_v$v.log
done
