// Shared device/host helpers for the dreamer4_b200 CUDA path (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define D4_WARP 32
#define D4_FULL 0xffffffffu

// fp32 machine epsilon: nn.RMSNorm(eps=None) resolves to torch.finfo(float32).eps
// (reference dreamer4/dreamer4.py:1906, 2089, 2822).
#define D4_RMS_EPS 1.1920928955078125e-07f
// F.normalize(p=2) eps (reference dreamer4/dreamer4.py:521-522)
#define D4_L2_EPS 1e-12f
// nn.LayerNorm default eps (x-mlps normed MLP)
#define D4_LN_EPS 1e-5f

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(D4_FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(D4_FULL, v, o));
    return v;
}
// fp32 -> nearest TF32 value (10 explicit mantissa bits), still an fp32 word.  The 3xTF32 GEMMs split both operands with
// ROUND-TO-NEAREST (x = hi + lo, hi = rna(x), lo = rna(x - hi)): the dropped lo*lo term and the representation error of lo
// are then unbiased and <= 2^-24 |x|; a truncating split leaves same-signed errors that add up coherently over K.
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
#ifdef D4_CUSIM      // host build for the CPU kernel simulator (tests/cusim): the same rounding in integer arithmetic
    r = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
#else
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
#endif
    return __uint_as_float(r);
}
// The same split for the warp-level mma.sync kernels (space_attn.cu, frame_attn_mma.cu), where it is paid per fragment element: ptxas
// expands cvt.rna.tf32.f32 into four instructions (add, Inf/NaN test, select, mask), nine per split.  Here: hi by integer rounding
// (finite operands; ties away from zero like cvt.rna), lo = x - hi exactly, plus half a TF32 ulp so that the tensor core's truncation
// of the 13 low operand bits rounds lo to nearest - four instructions.  hi and lo are bit patterns for the MMA operand registers.
__device__ __forceinline__ void tf32_split_mma(float x, uint32_t& hi, uint32_t& lo) {
    hi = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
    lo = __float_as_uint(x - __uint_as_float(hi)) + 0x1000u;
}
// tanh on |x| <= 0.75 (softclamped attention scores: |s| <= 37.5 at the default clamp of 50): x + x^3 P(x^2), P a degree-5
// least-squares fit on Chebyshev nodes; |error| <= 5.2e-8 over the interval evaluated in fp32 (under one ulp of the result near
// the interval's end), against ~35 instructions and a branch for tanhf.  Callers fall back to tanhf when a warp holds a larger argument.
constexpr float D4_TANH_POLY_MAX = 0.75f;
__device__ __forceinline__ float tanh_small_(float x) {
    const float u = x * x;
    float q = 0.0019145376281812787f;
    q = fmaf(q, u, -0.007940924726426601f);
    q = fmaf(q, u, 0.02161884494125843f);
    q = fmaf(q, u, -0.05393656715750694f);
    q = fmaf(q, u, 0.13333185017108917f);
    q = fmaf(q, u, -0.3333333134651184f);
    return fmaf(x * u, q, x);
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float siluf_(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float geluf_(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// row remap used by the GEMMs and the row-wise kernels: compact row m -> physical row
//   grp == 0 : m
//   else     : (m / grp) * gstride + goff + (m % grp)
struct RowMap {
    int grp, gstride, goff;
    __host__ __device__ __forceinline__ long long operator()(int m) const {
        if (grp == 0) return m;
        return (long long)(m / grp) * gstride + goff + (m % grp);
    }
};
static inline RowMap rowmap_identity() { RowMap r; r.grp = 0; r.gstride = 0; r.goff = 0; return r; }
static inline RowMap rowmap(int grp, int gstride, int goff) { RowMap r; r.grp = grp; r.gstride = gstride; r.goff = goff; return r; }

enum { D4_ACT_NONE = 0, D4_ACT_GLU_SILU = 1, D4_ACT_GLU_GELU = 2, D4_ACT_SILU = 3 };

// One linear layer  C = epi( rowscale[m] * (A @ W^T) + bias )  (+ residual)
// A (M,K) fp32 row-major with leading dim lda (rows through amap); W (N,K) fp32 (nn.Linear layout, leading dim ldw).
// act GLU_*: W rows are interleaved [x0,g0,x1,g1,...]; output has N/2 columns  x_j * act(g_j).
struct GemmArgs {
    const float* A; long long lda;
    const float* W; long long ldw;
    const float* W_lo;            // low words of W for the tf32x3 path (nullptr otherwise)
    float* C; long long ldc;
    int M, N, K;
    const float* bias;
    const float* row_scale;
    const float* residual; long long ldr;
    int act;
    RowMap amap, cmap;
    // operand layouts (exact-fp32 path only; used by the learn_from_experience backward GEMMs):
    //   transA: A is stored (K, M) — element (m,k) at A[k*lda + m];  transW: W is stored (K, N) — element (n,k) at W[k*ldw + n]
    int transA, transW;
    // CTA-pair tensor-core path only (gemm_tc3.cu, TMA epilogue, identity output rows):
    //   rs_mode 1: row_scale[m] holds the SUM OF SQUARES of A's row m; the epilogue scales by rsqrt(ss / K + eps) (the RMSNorm
    //              statistic, with K the normalised width) instead of reading a precomputed rstd
    //   ss_out   : the epilogue atomically adds sum_n C[m][n]^2 into ss_out[m] (feeds the next consumer's rs_mode 1)
    int rs_mode;
    float* ss_out;
};

static inline GemmArgs gemm_args(const float* A, long long lda, const float* W, long long ldw, float* C, long long ldc,
                                 int M, int N, int K) {
    GemmArgs g;
    g.A = A; g.lda = lda; g.W = W; g.ldw = ldw; g.W_lo = nullptr; g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.K = K;
    g.bias = nullptr; g.row_scale = nullptr; g.residual = nullptr; g.ldr = 0; g.act = D4_ACT_NONE;
    g.amap = rowmap_identity(); g.cmap = rowmap_identity(); g.transA = 0; g.transW = 0; g.rs_mode = 0; g.ss_out = nullptr;
    return g;
}

// every kernel launch of the library is counted (bench.py reports the count it saw in the timed region)
extern long long d4_launches_;
#define D4_COUNT_LAUNCH() (++d4_launches_)

#define D4_CUDA_OK(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return d4_fail_cuda(e__, #expr, __FILE__, __LINE__); } while (0)
int d4_fail_cuda(cudaError_t e, const char* what, const char* file, int line);
int d4_fail(const char* fmt, ...);
