"""CPU study: how much accuracy do the candidate operand splits of the tensor-core GEMM keep?  (No GPU needed: every product of
two split words is exact in fp32/fp64, so the splitting error can be emulated exactly; accumulation is done in fp64 to isolate
it from accumulation order.)

Splits of  D = A @ W^T  (A activations (M, K) fp32, W weights (N, K) fp32):

  tf32x3      current kernel: a_hi = trunc_tf32(a) (what the tensor core does to a raw fp32 operand), a_lo = rna_tf32(a - a_hi),
              w_hi = rna_tf32(w), w_lo = rna_tf32(w - w_hi);  D = a_hi.w_hi + a_lo.w_hi + a_hi.w_lo        (3 TF32 MMAs = 3 units)
  f16x3       the same three terms on kind::f16 (twice the TF32 rate): hi = fp16(x), lo = fp16(2^11 (x - hi)), the lo terms
              accumulated apart and scaled by 2^-11 (Ootomo & Yokota's error-corrected fp16 GEMM)             (3 F16 MMAs = 1.5 units)
  f16x3p      fp16 without the 2^11 trick and with ONE accumulator: rows of A and the whole of W are first multiplied by a power
              of two that brings their rms to ~1 (exact exponent shifts, undone by the epilogue's row scale), then
              hi = fp16(x), lo = fp16(x - hi);  D = hi.hi + lo.hi + hi.lo                                     (3 F16 MMAs = 1.5 units)
  tf32+f16x2  main term on TF32, the two corrections on kind::f16 as above                                     (2 units)
  bf16x3      hi / lo in bf16 (8-bit significands), for reference                                              (1.5 units)
  tf32x1      single TF32 pass, for reference                                                                  (1 unit)

fp16 carries the same 11-bit significand as TF32, so f16x3 differs from tf32x3 only through fp16's exponent range
(max 65504, normals down to 6.1e-5, subnormals to 6e-8) - which is what the `scale` sweep probes.

    python scripts/split_precision_study.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dreamer4_b200.packing import tf32_round  # noqa: E402


def tf32_trunc(x):
    return (x.contiguous().view(torch.int32) & -8192).view(torch.float32)


def f16(x):
    return x.to(torch.float16).to(torch.float32)


def bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def mm(a, w):
    return a.double() @ w.double().T


def variants(a, w):
    out = {}
    a_hi, w_hi = tf32_trunc(a), tf32_round(w)
    a_lo, w_lo = tf32_round(a - a_hi), tf32_round(w - w_hi)
    out['tf32x1'] = mm(a_hi, w_hi)
    out['tf32x3'] = mm(a_hi, w_hi) + mm(a_lo, w_hi) + mm(a_hi, w_lo)
    s = 2.0 ** 11
    ah, wh = f16(a), f16(w)
    al, wl = f16((a - ah) * s), f16((w - wh) * s)
    out['f16x3'] = mm(ah, wh) + (mm(al, wh) + mm(ah, wl)) / s
    pow2 = lambda x: 2.0 ** torch.round(torch.log2(x))
    p = 1.0 / pow2(a.pow(2).mean(dim=1, keepdim=True).sqrt().clamp(min=1e-30))          # per row of A
    q = 1.0 / pow2(w.pow(2).mean().sqrt())                                             # per weight matrix
    ap, wp = a * p, w * q
    aph, wph = f16(ap), f16(wp)
    out['f16x3p'] = (mm(aph, wph) + mm(f16(ap - aph), wph) + mm(aph, f16(wp - wph))) / (p.double() * float(q))
    al2, wl2 = f16((a - a_hi) * s), f16((w - w_hi) * s)
    out['tf32+f16x2'] = mm(a_hi, w_hi) + (mm(al2, f16(w_hi)) + mm(f16(a_hi), wl2)) / s
    bh, bwh = bf16(a), bf16(w)
    out['bf16x3'] = mm(bh, bwh) + mm(bf16(a - bh), bwh) + mm(bh, bf16(w - bwh))
    return out


def report(name, a, w):
    exact = mm(a, w)
    denom = a.double().abs() @ w.double().abs().T              # sum_k |a_k w_k|: the scale rounding errors live on
    fp32_ref = (a @ w.T).double()                              # what an fp32 FMA GEMM gives (accumulation error included)
    line = f'{name:<34s}'
    for k, v in variants(a, w).items():
        line += f' {k} {float(((v - exact).abs() / denom).max()):.1e}'
    line += f' | fp32-FMA {float(((fp32_ref - exact).abs() / denom).max()):.1e}'
    print(line)


def main():
    torch.manual_seed(0)
    print('max over outputs of |D_split - D_exact| / sum_k |a_k w_k|   (2^-24 = 6.0e-08 is one fp32 rounding)')
    for K, N in ((512, 1552), (1376, 512), (2048, 2048)):
        w = torch.randn(N, K) * K ** -0.5
        # residual-stream-like activations: unit-scale rows with a few large outliers
        a = torch.randn(256, K)
        a[:, :8] *= 30.
        report(f'K={K} N={N} rows~N(0,1)+outliers', a, w)
    K, N = 512, 512
    w = torch.randn(N, K) * K ** -0.5
    for scale in (1e-6, 1e-4, 1e-2, 1., 1e2, 1e4):
        report(f'K={K} activation scale {scale:g}', torch.randn(256, K) * scale, w)
    a = torch.randn(256, K) * 3e4                              # a few elements beyond fp16's 65504
    v = variants(a, w)
    print(f'fp16 overflow: {int((a.abs() > 65504).sum())} of {a.numel()} activations exceed 65504 at scale 3e4 ->',
          'f16x3 finite' if torch.isfinite(v['f16x3']).all() else 'f16x3 produces inf/nan',
          '; f16x3p (pre-scaled rows) finite' if torch.isfinite(v['f16x3p']).all() else '; f16x3p inf/nan')


if __name__ == '__main__':
    main()
