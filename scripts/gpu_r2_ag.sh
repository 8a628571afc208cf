#!/bin/bash
# wave-count N-tile choice (0.7) as the default: whole GPU suite, 256 / 2048 dreams against the pinned old behaviour is not possible any more -
# compare with call AF (default 20.0k / 32.7k; model 0.6: 19.8k / 33.25k); configs 2 and 3
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --batch 256 --horizon 16 --steps 3 --warmup 3 --no-cpu-baseline --no-weak --no-profile > gpurun_out/r2ag_b256.json 2> gpurun_out/r2ag_b256.err; python - <<PY
import json; d=json.loads(open('gpurun_out/r2ag_b256.json').read().strip().splitlines()[-1]); print('B=256', round(d['value'],1), d['ms_per_step'])
PY
timeout 600 python bench.py --horizon 16 --steps 3 --warmup 3 --no-cpu-baseline --no-weak --no-profile > gpurun_out/r2ag_b2048.json 2> gpurun_out/r2ag_b2048.err; python - <<PY
import json; d=json.loads(open('gpurun_out/r2ag_b2048.json').read().strip().splitlines()[-1]); print('B=2048', round(d['value'],1), d['ms_per_step'])
PY
for cfg in 2 3; do
timeout 600 python bench.py --workload config$cfg --steps 2 --warmup 3 --no-cpu-baseline --no-weak > gpurun_out/r2ag_config$cfg.json 2> gpurun_out/r2ag_config$cfg.err; python - <<PY
import json; d=json.loads(open('gpurun_out/r2ag_config$cfg.json').read().strip().splitlines()[-1]); print('config $cfg', round(d['value'],1), d['ms_per_step'], d['config']['workload'][:60])
PY
done
