#!/bin/bash
# register-resident LayerNorm+SiLU rows, row-dot kernel for the policy unembedding: whole GPU suite, head-kernel launch list, bench
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ln_act|rowdot|gemm_simt|lp_fused|l2s_fused' -s 100 -c 120 --csv --log-file gpurun_out/r2ac_launches.csv \
    python bench.py --horizon 6 --steps 1 --warmup 1 --no-cpu-baseline --no-profile --no-weak > gpurun_out/r2ac_launches.log 2>&1; echo "launch list rc=$?"
timeout 600 python bench.py --horizon 16 --steps 3 --warmup 3 --no-cpu-baseline --no-weak --no-profile > gpurun_out/r2ac_bench.json 2> gpurun_out/r2ac_bench.err; python - <<PY
import json; d=json.loads(open('gpurun_out/r2ac_bench.json').read().strip().splitlines()[-1]); print('bench h16', round(d['value'],1), d['ms_per_step'])
PY
