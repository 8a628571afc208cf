#!/bin/bash
# tokenizer f16x3 parity (the clamp in d4_tf_create is gone), frame graphs at 256 and 2048 dreams per GPU
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zy_tokenizer_gpu.py -x -q 2>&1 | tail -5
for rows in 4096 16384; do
  D4_GRAPH_MAX_ROWS=$rows timeout 600 python bench.py --batch 256 --horizon 16 --steps 3 --warmup 3 --no-cpu-baseline --no-profile --no-weak > gpurun_out/r2u_b256_rows$rows.json 2> gpurun_out/r2u_b256_rows$rows.err; python - <<PY
import json; d=json.loads(open('gpurun_out/r2u_b256_rows$rows.json').read().strip().splitlines()[-1]); print('B=256 graph rows $rows', round(d['value'],1), d['ms_per_step'], d.get('gpu_launches'))
PY
done
D4_GRAPH_MAX_ROWS=131072 timeout 600 python bench.py --horizon 16 --steps 3 --warmup 3 --no-cpu-baseline --no-profile --no-weak > gpurun_out/r2u_b2048_graph.json 2> gpurun_out/r2u_b2048_graph.err; python - <<PY
import json; d=json.loads(open('gpurun_out/r2u_b2048_graph.json').read().strip().splitlines()[-1]); print('B=2048 graphs', round(d['value'],1), d['ms_per_step'])
PY
timeout 600 python bench.py --horizon 16 --steps 3 --warmup 3 --no-cpu-baseline --no-profile --no-weak > gpurun_out/r2u_b2048_nograph.json 2> gpurun_out/r2u_b2048_nograph.err; python - <<PY
import json; d=json.loads(open('gpurun_out/r2u_b2048_nograph.json').read().strip().splitlines()[-1]); print('B=2048 direct', round(d['value'],1), d['ms_per_step'])
PY
