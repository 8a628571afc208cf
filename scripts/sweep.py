#!/usr/bin/env python
"""BASELINE.json configs[4]: frames/s over (dreams per GPU, horizon) at the config-4 model, one bench.py process per point.
Writes one JSON line per point (the bench line, trimmed) to stdout / the file given as argv[1]."""
import json
import subprocess
import sys

POINTS = [(256, 64), (2048, 8), (2048, 64), (8192, 32)] if len(sys.argv) > 2 and sys.argv[2] == "short" else [(256, 8), (256, 64), (1024, 64), (2048, 8), (2048, 64), (2048, 128), (8192, 8), (8192, 32)]
out = open(sys.argv[1], 'w') if len(sys.argv) > 1 else None
for b, h in POINTS:
    r = subprocess.run([sys.executable, 'bench.py', '--batch', str(b), '--horizon', str(h), '--steps', '2', '--warmup', '3', '--no-cpu-baseline'],
                       capture_output=True, text=True)
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ''
    try:
        d = json.loads(line)
        row = dict(dreams_per_gpu=b, horizon=h, frames_per_s=round(d['value'], 1), e2e_frames_per_s=round(d['e2e']['value'], 1), ms_per_step=round(d['ms_per_step'], 1),
                   k1_frac_hbm=round(d['roofline_attn']['frac'], 3), k1_gbs=round(d['roofline_attn']['achieved'], 1),
                   gemm_tflops=round(d['roofline_gemm']['achieved'], 1), gemm_frac_arith_bound=round(d['roofline_gemm'].get('frac_of_arith_bound', 0), 3),
                   share=d['kernel_class_share'], clocks=d['clocks'])
    except Exception as e:
        row = dict(dreams_per_gpu=b, horizon=h, error=str(e), stderr=r.stderr[-400:])
    s = json.dumps(row)
    print(s, flush=True)
    if out:
        out.write(s + '\n'); out.flush()
