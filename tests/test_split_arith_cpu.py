"""Arithmetic of the 4-instruction TF32 operand split used by the warp-level mma.sync kernels (csrc/common.cuh: tf32_split_mma), restated in
numpy integer / float32 arithmetic:  hi = (bits(x) + 0x1000) & ~0x1FFF,  lo_bits = bits(x - hi) + 0x1000, and the tensor core reads the top 19
bits of each operand word (it ignores the 13 low mantissa bits).  Claims checked: hi is x rounded to nearest TF32 (ties away from zero, as
cvt.rna.tf32.f32), x - hi is exact in fp32, the truncated lo is (x - hi) rounded to nearest TF32, and hi + lo reproduces x to 2^-21 |x|
(the bound DESIGN.md quotes for the 3-term products)."""
import numpy as np


def _bits(x):
    return x.view(np.uint32)


def _f32(b):
    return b.astype(np.uint32).view(np.float32)


def _trunc_tf32(b):
    return b & np.uint32(0xFFFFE000)


def _rna_tf32(x):
    """round-to-nearest TF32, ties away from zero, in float64 arithmetic (independent of the bit trick)"""
    x64 = x.astype(np.float64)
    m, e = np.frexp(x64)                       # x = m * 2^e, 0.5 <= |m| < 1
    scaled = m * 2.0 ** 11                     # 11 significand bits incl. the implicit one
    r = np.sign(scaled) * np.floor(np.abs(scaled) + 0.5)
    return (r / 2.0 ** 11 * 2.0 ** e).astype(np.float32)


def test_tf32_split_mma_matches_round_to_nearest_and_recombines():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(200000).astype(np.float32) * np.float32(s) for s in (1e-6, 1e-2, 1.0, 37.5, 3e4)])
    x = np.concatenate([x, np.float32([0.0, -0.0, 1.0, -1.0, 1.0 + 2.0 ** -11, 1.0 + 3 * 2.0 ** -12, np.float32(2.0) - np.float32(2.0 ** -23)])])
    hi_b = (_bits(x) + np.uint32(0x1000)) & np.uint32(0xFFFFE000)
    hi = _f32(hi_b)
    np.testing.assert_array_equal(hi, _rna_tf32(x))                                  # == cvt.rna.tf32.f32
    d = (x - hi).astype(np.float32)
    np.testing.assert_array_equal(d.astype(np.float64), x.astype(np.float64) - hi.astype(np.float64))     # the subtraction is exact
    lo_b = _bits(d) + np.uint32(0x1000)
    lo_seen = _f32(_trunc_tf32(lo_b))                                               # what the tensor core multiplies with
    np.testing.assert_array_equal(lo_seen, _rna_tf32(d))                            # truncation of (lo + half ulp) = round to nearest
    err = np.abs(hi.astype(np.float64) + lo_seen.astype(np.float64) - x.astype(np.float64))
    assert np.all(err <= np.abs(x.astype(np.float64)) * 2.0 ** -21)
    nz = x != 0
    assert float(np.max(err[nz] / np.abs(x[nz]))) <= 2.0 ** -22 + 1e-12             # measured bound: half a TF32 ulp of lo


def test_small_tanh_polynomial_stays_under_one_ulp_of_its_range():
    """tanh_small_ (common.cuh): x + x^3 P(x^2) on |x| <= 0.75, evaluated in float32 exactly as the kernel does (fma chain)."""
    c = [np.float32(v) for v in (0.0019145376281812787, -0.007940924726426601, 0.02161884494125843, -0.05393656715750694,
                                 0.13333185017108917, -0.3333333134651184)]
    x = np.linspace(-0.75, 0.75, 400001).astype(np.float32)
    u = (x * x).astype(np.float32)
    q = np.full_like(x, c[0])
    for k in c[1:]:
        q = (q.astype(np.float64) * u.astype(np.float64) + np.float64(k)).astype(np.float32)          # one rounding per fma
    y = ((x * u).astype(np.float32).astype(np.float64) * q.astype(np.float64) + x.astype(np.float64)).astype(np.float32)
    err = np.abs(y.astype(np.float64) - np.tanh(x.astype(np.float64)))
    assert float(err.max()) <= 6e-8                                                   # one float32 ulp at 0.5 .. 1 is 5.96e-8
