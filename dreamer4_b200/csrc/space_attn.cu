// K2, third version: the S x S (S <= 16) space attention of one frame (reference dreamer4.py:1968-2075 with the mask of :1781), one
// warp per (frame, kv head), head dim 64, with EVERY operand held in registers in the layout the warp-level tensor-core MMA wants -
// no shared memory, no __syncwarp, no staging pass.
//
// Why: the second version (attn.cu: space_attn_mma_kernel) staged Q / K / V in 10.5 KB of shared memory per warp and ran ~2,900
// dependent instructions per (frame, head) behind two exposed memory latencies: 145 us per launch at 2048 dreams against a 48 us
// HBM floor (316 MB in + out), 0.33 of its roofline and the largest share of the "small attention" class.
// The m16n8k8 fragments only fix WHICH 8 contraction indices a k-step covers and which lane holds which of them - the contraction
// index itself may be permuted freely as long as A and B use the same permutation, and so may the output columns.  Chosen so that
// every lane reads contiguous, coalesced 64-byte runs straight from global memory into its fragments:
//   g = lane / 4, t = lane % 4
//   Q K^T (contraction over the 64 head dims):  k-step s = 2 q + u, fragment index t <-> head dim 16 q + 4 t + 2 u,  t + 4 <-> the next one
//       => lane (g, t) holds the four 16-byte pieces at dims 16 q + 4 t of query rows g, g + 8 (A) and of key rows g, g + 8 (B);
//          the four lanes of a row read 64 contiguous bytes per instruction
//   P V (contraction over the 16 keys):  k-step s, fragment index t <-> key 8 s + 2 t,  t + 4 <-> key 8 s + 2 t + 1
//       => the score C-fragments ARE the probability A-fragments (no shuffle, no shared memory)
//       output tile n, fragment column g <-> head dim 4 g + (n & 3) + 32 (n >> 2)   (was 8 g + n until the L1 sector fix below)
//       => lane (g, t) holds dims [8 g, 8 g + 8) of value rows 2 t, 2 t + 1, 8 + 2 t, 9 + 2 t (B): 2 float4 per row, and ends up
//          with out[g | g + 8][16 t .. 16 t + 16): 4 float4 stores per row
// Products are 3-term TF32 splits (a_lo b_hi + a_hi b_lo + a_hi b_hi, operands rounded to nearest), fp32-accurate, as before.
// Fused exactly as before: key RMS-norm (gain folded into Q), softclamp, agent-key mask, softmax, value-residual lerp, belief
// projection out -= (out . vhat) vhat, per-head sigmoid gate.
#include <float.h>
#include <stdlib.h>
#include "kernels.h"

namespace {

constexpr int SPW = 4;          // warps per CTA

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) { tf32_split_mma(x, hi, lo); }
__device__ __forceinline__ void mma_3x(float (&c)[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4], const uint32_t (&bhi)[2], const uint32_t (&blo)[2]) {
    mma_tf32(c, alo, bhi);
    mma_tf32(c, ahi, blo);
    mma_tf32(c, ahi, bhi);
}
__device__ __forceinline__ float lerpf_(float a, float b, float w) { return a + w * (b - a); }
__device__ __forceinline__ void load16(float (&dst)[16], const float* p, bool ok) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 v = ok ? *reinterpret_cast<const float4*>(p + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
        dst[4 * q] = v.x; dst[4 * q + 1] = v.y; dst[4 * q + 2] = v.z; dst[4 * q + 3] = v.w;
    }
}
__device__ __forceinline__ void load8(float (&dst)[8], const float* p, bool ok) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const float4 v = ok ? *reinterpret_cast<const float4*>(p + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
        dst[4 * q] = v.x; dst[4 * q + 1] = v.y; dst[4 * q + 2] = v.z; dst[4 * q + 3] = v.w;
    }
}
// Sector-friendly fragment loads (round 2, last ncu capture: the L1 data path at 68 % with every sector requested twice - a lane's
// 16-byte pieces used to sit 64 / 32 bytes apart, so each instruction touched half-used sectors).  PIECES 16-byte pieces per lane,
// STRIDE floats apart: in ONE instruction the lanes of a row read adjacent pieces, i.e. whole 32-byte sectors.
template <int PIECES, int STRIDE>
__device__ __forceinline__ void load_pieces(float (&dst)[4 * PIECES], const float* p, bool ok) {
#pragma unroll
    for (int q = 0; q < PIECES; ++q) {
        const float4 v = ok ? *reinterpret_cast<const float4*>(p + STRIDE * q) : make_float4(0.f, 0.f, 0.f, 0.f);
        dst[4 * q] = v.x; dst[4 * q + 1] = v.y; dst[4 * q + 2] = v.z; dst[4 * q + 3] = v.w;
    }
}

// ONE_GROUP: one query head per kv head (no GQA): K / gain die after the scores, the compiler keeps everything in registers.
// MINB: resident CTAs per SM the register allocation is held to (4: 121 registers, no spills; 5: 96, 16 bytes; 6: 80, ~150 bytes).
template <bool ONE_GROUP, int MINB>
__global__ void __launch_bounds__(SPW * 32, MINB) space_attn_reg_kernel(SmallAttnArgs a) {
    constexpr int D = 64;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * SPW + warp;
    if (item >= (long long)a.nb * a.hkv) return;
    const int b = (int)(item / a.hkv), hk = (int)(item % a.hkv);
    const int S = a.n;
    const int g = lane >> 2, t = lane & 3;
    const float sqrt_d = 8.f;

    // ---- keys of rows g, g + 8, head dims [16 t, 16 t + 16); their l2 norms (summed over the four lanes t of a row)
    float kf[2][16];
#pragma unroll
    for (int h = 0; h < 2; ++h) load_pieces<4, 16>(kf[h], a.k + b * a.k_sb + (long long)(g + 8 * h) * a.k_sj + (long long)hk * D + 4 * t, g + 8 * h < S);
    float kinv_own[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float ss = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) ss = fmaf(kf[h][c], kf[h][c], ss);
        ss += __shfl_xor_sync(D4_FULL, ss, 1); ss += __shfl_xor_sync(D4_FULL, ss, 2);
        kinv_own[h] = 1.f / fmaxf(sqrtf(ss), D4_L2_EPS);
    }
    // 1 / |k_j| for the keys this lane's score fragments see: j = 8 n + 2 t + e, held by the lanes of row g' = 2 t + e
    float kinv[2][2];
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int e = 0; e < 2; ++e) kinv[n][e] = __shfl_sync(D4_FULL, kinv_own[n], 4 * (2 * t + e));
    // key-norm gain (gamma + 1) sqrt(d) of this lane's 16 head dims: folded into the queries
    float gain[16];
    load_pieces<4, 16>(gain, a.k_gamma + hk * D + 4 * t, true);
#pragma unroll
    for (int c = 0; c < 16; ++c) gain[c] = (gain[c] + 1.f) * sqrt_d;

    float vb[4][8], vr[2][16];
    float vinv[2] = {0.f, 0.f};
    // ---- values (+ value-residual lerp) in the two layouts they are consumed in, fetched once (first query group) and only when
    // the operands of the previous phase are dead - the kernel is register-bound (128 per thread at 16 warps per SM):
    //   vb: rows 2t, 2t+1, 8+2t, 9+2t, head dims [8 g, 8 g + 8)  - B operand of P V
    //   vr: rows g, g + 8, head dims [16 t, 16 t + 16)           - the query token's own value for the belief projection
    auto fetch_vb = [&]() {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int j = 8 * (r >> 1) + 2 * t + (r & 1);
            load_pieces<2, 32>(vb[r], a.v + b * a.v_sb + (long long)j * a.v_sj + (long long)hk * D + 4 * g, j < S);
        }
    };
    auto lerp_vb = [&]() {
        if (!a.v0) return;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int j = 8 * (r >> 1) + 2 * t + (r & 1);
            if (j < S) {
                float r0[8];
                load_pieces<2, 32>(r0, a.v0 + b * a.v0_sb + (long long)j * a.v0_sj + (long long)hk * D + 4 * g, true);
                const float w = sigmoidf_(a.mix[b * a.mix_sb + (long long)j * a.mix_sj + hk]);
#pragma unroll
                for (int c = 0; c < 8; ++c) vb[r][c] = lerpf_(vb[r][c], r0[c], w);
            }
        }
    };
    auto fetch_vr = [&]() {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float* vp = a.v + b * a.v_sb + (long long)(g + 8 * h) * a.v_sj + (long long)hk * D + 8 * t;
            float lo8[8], hi8[8];
            load8(lo8, vp, g + 8 * h < S); load8(hi8, vp + 32, g + 8 * h < S);
#pragma unroll
            for (int c = 0; c < 8; ++c) { vr[h][c] = lo8[c]; vr[h][8 + c] = hi8[c]; }
        }
    };
    auto lerp_vr = [&]() {
        if (a.v0) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = g + 8 * h;
                if (i < S) {
                    float r0[16];
                    {
                        const float* rp = a.v0 + b * a.v0_sb + (long long)i * a.v0_sj + (long long)hk * D + 8 * t;
                        float lo8[8], hi8[8];
                        load8(lo8, rp, true); load8(hi8, rp + 32, true);
#pragma unroll
                        for (int c = 0; c < 8; ++c) { r0[c] = lo8[c]; r0[8 + c] = hi8[c]; }
                    }
                    const float w = sigmoidf_(a.mix[b * a.mix_sb + (long long)i * a.mix_sj + hk]);
#pragma unroll
                    for (int c = 0; c < 16; ++c) vr[h][c] = lerpf_(vr[h][c], r0[c], w);
                }
            }
        }
        if (a.belief) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float ss = 0.f;
#pragma unroll
                for (int c = 0; c < 16; ++c) ss = fmaf(vr[h][c], vr[h][c], ss);
                ss += __shfl_xor_sync(D4_FULL, ss, 1); ss += __shfl_xor_sync(D4_FULL, ss, 2);
                vinv[h] = 1.f / fmaxf(sqrtf(ss), D4_L2_EPS);
            }
        }
    };

    const int groups = ONE_GROUP ? 1 : a.g;
#pragma unroll 1
    for (int gi = 0; gi < groups; ++gi) {
        const int hq = hk * groups + gi;
        // ---- queries of rows g, g + 8 (gain folded in); scores: 16 queries x 16 keys = two 8-key tiles, 8 k-steps
        float qf[2][16];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            load_pieces<4, 16>(qf[h], a.q + b * a.q_sb + (long long)(g + 8 * h) * a.q_si + (long long)hq * D + 4 * t, g + 8 * h < S);
#pragma unroll
            for (int c = 0; c < 16; ++c) qf[h][c] *= gain[c];
        }
        float sc[2][4];
#pragma unroll
        for (int n = 0; n < 2; ++n)
#pragma unroll
            for (int r = 0; r < 4; ++r) sc[n][r] = 0.f;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            uint32_t ahi[4], alo[4];
            split_tf32(qf[0][2 * s], ahi[0], alo[0]);
            split_tf32(qf[1][2 * s], ahi[1], alo[1]);
            split_tf32(qf[0][2 * s + 1], ahi[2], alo[2]);
            split_tf32(qf[1][2 * s + 1], ahi[3], alo[3]);
#pragma unroll
            for (int n = 0; n < 2; ++n) {
                uint32_t bhi[2], blo[2];
                split_tf32(kf[n][2 * s], bhi[0], blo[0]);
                split_tf32(kf[n][2 * s + 1], bhi[1], blo[1]);
                mma_3x(sc[n], ahi, alo, bhi, blo);
            }
        }
        if (gi == 0) fetch_vb();                      // in flight under the softmax
        // scale by the key norm, softclamp, mask, softmax over the keys of each query row (rows g: regs 0, 1; g + 8: regs 2, 3)
        if (a.softclamp > 0.f) {          // tanh(s / c) c  (reference dreamer4.py:1723-1724)
            const float pre = a.scale * (1.f / a.softclamp);
            float amax = 0.f;
#pragma unroll
            for (int n = 0; n < 2; ++n)
#pragma unroll
                for (int r = 0; r < 4; ++r) { sc[n][r] = sc[n][r] * kinv[n][r & 1] * pre; amax = fmaxf(amax, fabsf(sc[n][r])); }
            if (__all_sync(D4_FULL, amax <= D4_TANH_POLY_MAX)) {
#pragma unroll
                for (int n = 0; n < 2; ++n)
#pragma unroll
                    for (int r = 0; r < 4; ++r) sc[n][r] = tanh_small_(sc[n][r]) * a.softclamp;
            } else {
#pragma unroll
                for (int n = 0; n < 2; ++n)
#pragma unroll
                    for (int r = 0; r < 4; ++r) sc[n][r] = tanhf(sc[n][r]) * a.softclamp;
            }
        } else {
#pragma unroll
            for (int n = 0; n < 2; ++n)
#pragma unroll
                for (int r = 0; r < 4; ++r) sc[n][r] = sc[n][r] * kinv[n][r & 1] * a.scale;
        }
#pragma unroll
        for (int n = 0; n < 2; ++n)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int i = g + ((r & 2) ? 8 : 0), j = n * 8 + 2 * t + (r & 1);
                if (a.mask_agent && i < S - 1 && j == S - 1) sc[n][r] = -FLT_MAX;
                if (j >= S) sc[n][r] = -INFINITY;
            }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float mx = fmaxf(fmaxf(sc[0][2 * half], sc[0][2 * half + 1]), fmaxf(sc[1][2 * half], sc[1][2 * half + 1]));
            mx = fmaxf(mx, __shfl_xor_sync(D4_FULL, mx, 1)); mx = fmaxf(mx, __shfl_xor_sync(D4_FULL, mx, 2));
            float e[4], sum = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) { e[q] = __expf(sc[q >> 1][2 * half + (q & 1)] - mx); sum += e[q]; }
            sum += __shfl_xor_sync(D4_FULL, sum, 1); sum += __shfl_xor_sync(D4_FULL, sum, 2);
            const float inv = 1.f / sum;
#pragma unroll
            for (int q = 0; q < 4; ++q) sc[q >> 1][2 * half + (q & 1)] = e[q] * inv;
        }
        if (gi == 0) { lerp_vb(); fetch_vr(); }       // vr in flight under the P V products
        // ---- out = P V: the probability C-fragments are the A-fragments of k-step s = tile n (a0 c0, a1 c2, a2 c1, a3 c3)
        uint32_t phi[2][4], plo[2][4];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            split_tf32(sc[s][0], phi[s][0], plo[s][0]);
            split_tf32(sc[s][2], phi[s][1], plo[s][1]);
            split_tf32(sc[s][1], phi[s][2], plo[s][2]);
            split_tf32(sc[s][3], phi[s][3], plo[s][3]);
        }
        float o[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) {
#pragma unroll
            for (int r = 0; r < 4; ++r) o[n][r] = 0.f;
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                uint32_t bhi[2], blo[2];
                split_tf32(vb[2 * s][n], bhi[0], blo[0]);
                split_tf32(vb[2 * s + 1][n], bhi[1], blo[1]);
                mma_3x(o[n], phi[s], plo[s], bhi, blo);
            }
        }
        // lane (g, t) now holds out[g + 8 h][16 t + n] = o[n][2 h] and out[g + 8 h][16 t + 8 + n] = o[n][2 h + 1]
        if (gi == 0) lerp_vr();
        // ---- belief projection, head gate, store
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = g + 8 * h;
            float row[16];
#pragma unroll
            for (int n = 0; n < 4; ++n) {          // tile n, columns 2 t, 2 t + 1 <-> head dims 8 t + n, 8 t + 4 + n (n < 4), + 32 for tiles 4..7
                row[n] = o[n][2 * h]; row[4 + n] = o[n][2 * h + 1];
                row[8 + n] = o[4 + n][2 * h]; row[12 + n] = o[4 + n][2 * h + 1];
            }
            if (a.belief) {
                float dot = 0.f;
#pragma unroll
                for (int c = 0; c < 16; ++c) dot = fmaf(row[c], vr[h][c] * vinv[h], dot);
                dot += __shfl_xor_sync(D4_FULL, dot, 1); dot += __shfl_xor_sync(D4_FULL, dot, 2);
#pragma unroll
                for (int c = 0; c < 16; ++c) row[c] -= dot * (vr[h][c] * vinv[h]);
            }
            if (i < S) {
                const float gate = a.gate ? sigmoidf_(a.gate[b * a.gate_sb + (long long)i * a.gate_si + hq]) : 1.f;
                float* op = a.out + b * a.out_sb + (long long)i * a.out_si + (long long)hq * D + 8 * t;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<float4*>(op + 4 * (q & 1) + 32 * (q >> 1)) = make_float4(row[4 * q] * gate, row[4 * q + 1] * gate, row[4 * q + 2] * gate, row[4 * q + 3] * gate);
            }
        }
    }
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

int d4_space_attn_reg_ok(const SmallAttnArgs& a) {
    // one query head per kv head: with grouped queries K / V must stay live across the groups and the register budget (128 at 16 warps
    // per SM) spills - those configurations keep the shared-memory staged kernel of attn.cu
    return a.nq == a.n && a.n >= 1 && a.n <= 16 && a.d == 64 && a.g == 1 &&
           ((a.q_sb | a.q_si | a.k_sb | a.k_sj | a.v_sb | a.v_sj | a.v0_sb | a.v0_sj | a.out_sb | a.out_si) & 3) == 0 &&
           al16(a.q) && al16(a.k) && al16(a.v) && (!a.v0 || (al16(a.v0) && a.mix)) && al16(a.k_gamma) && al16(a.out) ? 1 : 0;
}

int d4_space_attn_reg(const SmallAttnArgs& a, cudaStream_t s) {
    if (!d4_space_attn_reg_ok(a)) return d4_fail("space_attn_reg: shape / alignment not supported");
    const long long items = (long long)a.nb * a.hkv;
    static int minb = 0;
    if (!minb) { const char* v = getenv("D4_SPACE_MINB"); minb = v ? atoi(v) : 4; }
    const unsigned grid = (unsigned)((items + SPW - 1) / SPW);
    if (minb == 6) space_attn_reg_kernel<true, 6><<<grid, SPW * 32, 0, s>>>(a);
    else if (minb == 5) space_attn_reg_kernel<true, 5><<<grid, SPW * 32, 0, s>>>(a);
    else space_attn_reg_kernel<true, 4><<<grid, SPW * 32, 0, s>>>(a);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError());
    return 0;
}
