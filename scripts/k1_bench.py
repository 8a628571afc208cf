#!/usr/bin/env python
"""K1 (time-decode attention over the KV cache) timed alone through the C-ABI at the config-4 shape:
M = dreams * 15 tokens, 8 heads x 64, for a sweep of context lengths t.  Prints achieved algorithmic GB/s
(M*h*d*4*(2t+4) bytes, +2 rows when appending) against the measured HBM peak."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from dreamer4_b200 import _lib as L  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--dreams', type=int, default=2048)
ap.add_argument('--tmax', type=int, default=64)
ap.add_argument('--variant', type=int, default=1)
ap.add_argument('--iters', type=int, default=5)
ap.add_argument('--ts', default='0,1,2,4,8,16,24,32,40,48,56,63')
args = ap.parse_args()
lib = L.load()
h = hq = 8; d = 64
M = args.dreams * 15
Dq = Dkv = h * d
ld = (Dq + 2 * Dkv + hq + h + 3) // 4 * 4
torch.manual_seed(0)
qkvgm = torch.randn(M, ld, device='cuda')
v0 = torch.randn(M, Dkv, device='cuda')
k_gamma = torch.zeros(h, d, device='cuda')
inv_freq = (1.0 / (10000. ** (torch.arange(0, d, 2).float() / d))).cuda()
kc = torch.randn(M, h, args.tmax, d, device='cuda')
vc = torch.randn(M, h, args.tmax, d, device='cuda')
out = torch.empty(M, Dq, device='cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
peak = 6550.4
pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
if os.path.exists(pk):
    peak = json.load(open(pk)).get('hbm_gbs', peak)
stream = torch.cuda.current_stream().cuda_stream
rows = []
for t in [int(x) for x in args.ts.split(',')]:
    if t >= args.tmax:
        continue
    best = 1e9
    for it in range(args.iters + 1):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(lib.d4_time_attn_decode(M, h, hq, d, t, args.tmax, L.ptr(qkvgm), ld, L.ptr(v0), L.ptr(k_gamma), L.ptr(inv_freq), L.ptr(kc), L.ptr(vc),
                                        L.ptr(out), 50.0, 0, args.variant, stream))
        e1.record()
        torch.cuda.synchronize()
        if it:
            best = min(best, e0.elapsed_time(e1))
    byts = M * h * d * 4 * (2 * t + 4)
    gbs = byts / (best * 1e-3) / 1e9
    rows.append(dict(t=t, us=best * 1e3, gbs=gbs, frac=gbs / peak))
    print(f't={t:3d}  {best * 1e3:8.1f} us  {gbs:7.1f} GB/s  {gbs / peak:5.3f} of measured HBM peak', flush=True)
print(json.dumps(dict(kernel='K1', variant=args.variant, M=M, peak_gbs=peak, rows=rows)))
