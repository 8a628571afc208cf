"""Do the tensor-core GEMMs give the same bits for a row whatever else is in the launch?  Rows 0..m-1 of an (M, K) product computed alone
and as part of a taller launch (other tile counts -> other N-tile widths), f16x3 and tf32x3."""
import ctypes as C
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from dreamer4_b200 import _lib as L  # noqa: E402
from dreamer4_b200.packing import f16_split, tf32_split  # noqa: E402

lib = L.load()
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
torch.manual_seed(0)
for N, K in ((512, 512), (256, 512), (512, 256)):
    W = torch.randn(N, K, device='cuda') / math.sqrt(K)
    hi16, lo16, inv_q = f16_split(W)
    hi32, lo32 = tf32_split(W)
    A = torch.randn(10200, K, device='cuda')
    outs = {}
    for prec in (3, 2):
        for M in (160, 2720, 10200):
            Cc = torch.zeros(M, N, device='cuda')
            rs = torch.full((M,), inv_q if prec == 3 else 1., device='cuda')
            if prec == 3:
                L.check(lib.d4_linear(3, M, N, K, L.ptr(A), K, L.ptr(hi16), K, L.ptr(lo16), None, L.ptr(rs), None, 0, 0, L.ptr(Cc), N, stream))
            else:
                L.check(lib.d4_linear(2, M, N, K, L.ptr(A), K, L.ptr(hi32), K, L.ptr(lo32), None, L.ptr(rs), None, 0, 0, L.ptr(Cc), N, stream))
            torch.cuda.synchronize()
            outs[(prec, M)] = Cc
        a, b, c = outs[(prec, 160)], outs[(prec, 2720)], outs[(prec, 10200)]
        print(f'N={N} K={K} prec={"f16x3" if prec == 3 else "tf32x3"}: rows 0..159 of M=160 vs 2720 equal: {torch.equal(a, b[:160])}, vs 10200: {torch.equal(a, c[:160])}; '
              f'2720 vs 10200: {torch.equal(b, c[:2720])}; max diff {float((b - c[:2720]).abs().max()):.3e}')
