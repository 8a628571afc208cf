"""Stand-alone timing of the tensor-core GEMM kernels on the layer shapes of the config-4 pass (through the C-ABI d4_linear):
tf32x3 (gemm_tc3.cu) against f16x3 (gemm_f16.cu), with the f16 kernel's ablation switches (d4_debug_set("gemm_f16", bits):
1 no epilogue, 2 no operand split, 4 no MMA; 7 = the TMA loads and the barrier protocol alone, i.e. the data-movement floor).  One JSON line per
(shape, mode), appended to gpurun_out/gemm_bench.jsonl.  The first f16x3 run of each shape is checked against fp64.

    python scripts/gemm_bench.py [--rows 30720,3840] [--reps 20] [--modes tf32x3,f16x3,f16x3_noepi,f16x3_nosplit,f16x3_nomma,f16x3_only_tma]
"""
import argparse
import ctypes as C
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from dreamer4_b200 import _lib as L  # noqa: E402
from dreamer4_b200.packing import f16_split, tf32_split  # noqa: E402

# (name, rows multiplier n (M = n * rows), N, K, act, residual, bias, rs_mode)
SHAPES = [
    ('qkv+gates (N=1552,K=512)', 1, 1552, 512, 0, False, True, 1),
    ('attn out (N=512,K=512,res)', 1, 512, 512, 0, True, False, 0),
    ('ff in (N=2730,K=512,GLU)', 1, 2730, 512, 1, False, True, 1),
    ('ff out (N=512,K=1376,res)', 1, 512, 1376, 0, True, True, 0),
    ('pool q (N=256,K=512)', 1, 256, 512, 0, False, False, 1),
    ('pool kv n=9 (N=512,K=512)', 9, 512, 512, 0, False, False, 1),
    ('pool out (N=512,K=256,res)', 1, 512, 256, 0, True, False, 0),
]
MODES = dict(tf32x3=(2, 0), f16x3=(3, 0), f16x3_noepi=(3, 1), f16x3_nosplit=(3, 2), f16x3_nomma=(3, 4), f16x3_only_tma=(3, 7), f16x3_only_tma_sameA=(3, 15), f16x3_only_tma_sameW=(3, 23), f16x3_only_tma_sameAW=(3, 31),
             f16x3_res_late=(3, 64), f16x3_nostore=(3, 128))          # residual chunk loads issued after the chunk's store, as before round 2's last change


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--rows', default='30720,3840')
    ap.add_argument('--reps', type=int, default=20)
    ap.add_argument('--modes', default='tf32x3,f16x3,f16x3_noepi,f16x3_nosplit,f16x3_nomma,f16x3_only_tma')
    args = ap.parse_args()
    lib = L.load()
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    out = open(os.path.join(ROOT, 'gpurun_out', 'gemm_bench.jsonl'), 'a')
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    for rows in [int(r) for r in args.rows.split(',')]:
        for name, mult, N, K, act, has_res, has_bias, rs_mode in SHAPES:
            M = rows * mult
            torch.manual_seed(1)
            A = torch.randn(M, K, device='cuda')
            W = (torch.randn(N, K, device='cuda') / math.sqrt(K))
            ldw16 = (K + 7) // 8 * 8
            hi16, lo16, inv_q = f16_split(W)
            if ldw16 != K:
                pad = lambda t: torch.nn.functional.pad(t, (0, ldw16 - K)).contiguous()
                hi16, lo16 = pad(hi16), pad(lo16)
            hi32, lo32 = tf32_split(W)
            bias = torch.randn(N, device='cuda') if has_bias else None
            nout = N // 2 if act else N
            res = torch.randn(M, nout, device='cuda') if has_res else None
            rs = (torch.rand(M, device='cuda') + 0.5)
            rsq = rs * inv_q
            ldc = (nout + 31) // 32 * 32
            Cc = torch.zeros(M, ldc, device='cuda')
            flops = 2.0 * M * N * K
            checked = False
            for mode in args.modes.split(','):
                prec, dbg = MODES[mode]
                L.check(lib.d4_debug_set(b'gemm_f16', dbg))

                def run():
                    if prec == 3:
                        return lib.d4_linear(3, M, N, K, L.ptr(A), K, L.ptr(hi16), ldw16, L.ptr(lo16), L.ptr(bias), L.ptr(rsq), L.ptr(res),
                                             nout, act, L.ptr(Cc), ldc, stream)
                    return lib.d4_linear(2, M, N, K, L.ptr(A), K, L.ptr(hi32), K, L.ptr(lo32), L.ptr(bias), L.ptr(rs), L.ptr(res), nout, act,
                                         L.ptr(Cc), ldc, stream)
                for _ in range(3):
                    L.check(run())
                torch.cuda.synchronize()
                err = None
                if prec == 3 and dbg == 0 and not checked:
                    ref = (A.double() @ W.double().T) * rs.double()[:, None] + (bias.double() if has_bias else 0.)
                    if act:
                        ref = ref[:, 0::2] * torch.nn.functional.silu(ref[:, 1::2])
                    elif has_res:
                        ref = ref + res.double()
                    err = float((Cc[:, :nout].double() - ref).abs().max() / ref.abs().max())
                    checked = True
                    del ref
                ms = []
                for _ in range(args.reps):
                    flush.zero_()                                   # L2 flush between timed launches
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    L.check(run())
                    e1.record()
                    torch.cuda.synchronize()
                    ms.append(e0.elapsed_time(e1))
                ms.sort()
                med = ms[len(ms) // 2]
                line = dict(shape=name, M=M, N=N, K=K, mode=mode, us=round(med * 1e3, 1), us_min=round(ms[0] * 1e3, 1),
                            tflops=round(flops / (med * 1e-3) / 1e12, 1), rel_err_vs_fp64=err)
                print(json.dumps(line), flush=True)
                out.write(json.dumps(line) + '\n')
                out.flush()
    L.check(lib.d4_debug_set(b'gemm_f16', 0))


if __name__ == '__main__':
    main()
