#!/bin/bash
# rollout head MLPs on the fp16 split GEMM in f16x3 mode: whole GPU suite, A/B is against call AC (32.70k at horizon 16), 256-dream point
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --horizon 16 --steps 3 --warmup 3 --no-cpu-baseline --no-weak --no-profile > gpurun_out/r2ad_bench.json 2> gpurun_out/r2ad_bench.err; python - <<PY
import json; d=json.loads(open('gpurun_out/r2ad_bench.json').read().strip().splitlines()[-1]); print('bench h16', round(d['value'],1), d['ms_per_step'], d.get('losses'))
PY
timeout 600 python bench.py --batch 256 --horizon 16 --steps 3 --warmup 3 --no-cpu-baseline --no-weak --no-profile > gpurun_out/r2ad_b256.json 2> gpurun_out/r2ad_b256.err; python - <<PY
import json; d=json.loads(open('gpurun_out/r2ad_b256.json').read().strip().splitlines()[-1]); print('B=256 h16', round(d['value'],1), d['ms_per_step'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm_tc3|gemm_f16x3_kernel<128>|ln_act' -s 60 -c 60 --csv --log-file gpurun_out/r2ad_launches.csv \
    python bench.py --horizon 4 --steps 1 --warmup 1 --no-cpu-baseline --no-profile --no-weak > gpurun_out/r2ad_launches.log 2>&1; echo "launch list rc=$?"
