"""EXPERIMENTAL kernel, first hardware run pending: the fp16 3-term split GEMM (dreamer4_b200/csrc/gemm_f16.cu) through
d4_linear(precision = D4_PREC_F16X3), against fp64, held to the SAME tolerance as the 3xTF32 kernel in
tests/test_gpu_parity.py::test_linear_tcgen05 - the point of the kernel is that accuracy at half the tensor-core cost
(scripts/split_precision_study.py).  The engine does not dispatch to it yet.

STATUS: written after round 1's GPU budget was spent - compiled for sm_100a, never executed.  It therefore runs only when asked
for (D4_EXPERIMENTAL=1): a never-run kernel that traps (its barrier waits trap instead of hanging) would poison the CUDA context
of the whole pytest process, which the regular -m gpu suite must not risk; the file also sorts last (test_zz*) for that reason.

    D4_EXPERIMENTAL=1 python -m pytest tests/test_zz_gemm_f16_gpu.py -q -m gpu
"""
import os
import ctypes as C
import math

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get('D4_EXPERIMENTAL') != '1', reason='experimental kernel: set D4_EXPERIMENTAL=1')]

D4_PREC_F16X3 = 3


def pack_f16(W):
    """hi / lo fp16 words of q W (q the power of two that brings rms(q W) to ~1), leading dimension padded to 8, and 1 / q."""
    q = 2.0 ** torch.round(torch.log2(1.0 / W.pow(2).mean().sqrt()))
    Wq = W * q
    N, K = W.shape
    ld = (K + 7) // 8 * 8
    hi = torch.zeros(N, ld, dtype=torch.float16, device=W.device)
    lo = torch.zeros(N, ld, dtype=torch.float16, device=W.device)
    hi[:, :K] = Wq.half()
    lo[:, :K] = (Wq - hi[:, :K].float()).half()
    return hi, lo, ld, float(1.0 / q)


@pytest.mark.parametrize('M,N,K,act', [(256, 256, 64, 0), (128, 128, 512, 0), (300, 200, 96, 0), (1000, 1552, 512, 0), (515, 2730, 512, 1),
                                      (700, 512, 1376, 0), (4096, 260, 512, 0), (130, 170, 64, 2), (3840, 1024, 32, 0), (2048, 2048, 2048, 0)])
def test_linear_f16x3(M, N, K, act):
    from dreamer4_b200 import _lib as L
    lib = L.load()
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K).cuda()
    A[:, :4] *= 30.                                   # a few outlier columns, like a residual stream
    W = (torch.randn(N, K) / math.sqrt(K)).cuda()
    hi, lo, ldw, inv_q = pack_f16(W)
    bias, rs = torch.randn(N).cuda(), torch.rand(M).cuda() + 0.5
    nout = N // 2 if act else N
    res = torch.randn(M, nout).cuda() if not act else None
    Cc = torch.full((M, nout), float('nan')).cuda()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    L.check(lib.d4_linear(D4_PREC_F16X3, M, N, K, L.ptr(A), K, L.ptr(hi), ldw, L.ptr(lo), L.ptr(bias), L.ptr(rs * inv_q), L.ptr(res),
                          nout, act, L.ptr(Cc), nout, stream))
    torch.cuda.synchronize()
    ref = (A.double() @ W.double().T) * rs.double()[:, None] + bias.double()
    if act:
        x, g = ref[:, 0::2], ref[:, 1::2]
        ref = x * (torch.nn.functional.silu(g) if act == 1 else torch.nn.functional.gelu(g))
    else:
        ref = ref + res.double()
    assert not torch.isnan(Cc).any()
    tol = 1e-5 * max(1.0, K / 256)
    err = (Cc.double() - ref).abs().max().item()
    assert err < tol * max(1.0, ref.abs().max().item()), f'max abs err {err}'
