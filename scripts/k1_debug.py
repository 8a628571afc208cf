import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from dreamer4_b200 import _lib as L
lib = L.load()
h = hq = 8; d = 64; M = int(os.environ.get('KM', '150')); Tmax = 256
Dq = Dkv = h * d; ld = (Dq + 2 * Dkv + hq + h + 3) // 4 * 4
torch.manual_seed(0)
qkvgm = torch.randn(M, ld, device='cuda'); v0 = torch.randn(M, Dkv, device='cuda'); k_gamma = torch.zeros(h, d, device='cuda')
inv_freq = (1.0 / (10000. ** (torch.arange(0, d, 2).float() / d))).cuda()
kc = torch.randn(M, h, Tmax, d, device='cuda'); vc = torch.randn(M, h, Tmax, d, device='cuda')
s = torch.cuda.current_stream().cuda_stream
for t in [int(x) for x in sys.argv[1].split(',')]:
    outs = []
    for variant in (0, 1):
        out = torch.zeros(M, Dq, device='cuda')
        L.check(lib.d4_time_attn_decode(M, h, hq, d, t, Tmax, L.ptr(qkvgm), ld, L.ptr(v0), L.ptr(k_gamma), L.ptr(inv_freq), L.ptr(kc), L.ptr(vc), L.ptr(out), 50.0, 0, variant, s))
        torch.cuda.synchronize(); outs.append(out)
    err = (outs[0] - outs[1]).abs()
    print(f't={t}: max err {err.max().item():.3e}; per-head max err', [f'{e:.1e}' for e in err.view(M, h, d).amax(dim=(0, 2)).tolist()], 'rows bad', (err.amax(dim=1) > 1e-4).sum().item(), flush=True)
