#!/bin/bash
# persistent space -> latent pool kernel + cheap tf32 split in the latent -> space pool: parity, A/B, launch list
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_horizon_parity_gpu.py tests/test_gpu_parity.py tests/test_fullsize_properties_gpu.py -x -q 2>&1 | tail -4
for v in 1 2 1 2; do
D4_LP_V=$v timeout 600 python bench.py --horizon 16 --steps 3 --warmup 3 --no-cpu-baseline --no-weak --no-profile > gpurun_out/r2aa_bench_$v.json 2> gpurun_out/r2aa_bench_$v.err; python - <<PY
import json; d=json.loads(open('gpurun_out/r2aa_bench_$v.json').read().strip().splitlines()[-1]); print('lp_v $v', round(d['value'],1), d['ms_per_step'])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'fused|space_attn|pool_attn' -s 200 -c 200 --csv --log-file gpurun_out/r2aa_launches.csv \
    python bench.py --horizon 6 --steps 1 --warmup 1 --no-cpu-baseline --no-profile --no-weak > gpurun_out/r2aa_launches.log 2>&1; echo "launch list rc=$?"
