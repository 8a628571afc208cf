#!/bin/bash
# occupancy variants of the register space-attention kernel after its instruction diet; fused RMS statistics on / off
set -u
mkdir -p gpurun_out
run() { env "$@" timeout 600 python bench.py --horizon 16 --steps 3 --warmup 3 --no-cpu-baseline --no-weak --no-profile > gpurun_out/r2ah.json 2> gpurun_out/r2ah.err; python - "$*" <<PY
import json,sys; d=json.loads(open('gpurun_out/r2ah.json').read().strip().splitlines()[-1]); print(sys.argv[1], round(d['value'],1), d['ms_per_step'])
PY
}
run D4_SPACE_MINB=4
run D4_SPACE_MINB=5
run D4_SPACE_MINB=6
run D4_SPACE_MINB=4
run D4_SPACE_MINB=5
run D4_FUSE_SS=0
