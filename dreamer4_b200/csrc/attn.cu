// Attention kernels of the imagination pass.
//
//  small_attn_kernel : every attention on the path whose key set fits in shared memory (<= 64 keys):
//      space attention over the S tokens of a frame (reference dreamer4.py:1968-2075 + naive_attend 1683-1756
//      with softclamp, agent-token mask 1769-1783, value-residual lerp 2005-2012, belief projection 2049-2054),
//      the learned-query latent<->spatial pools (2179-2210), the attention-residual pools over layer hiddens
//      (2143-2177) and the final agent cross-attention (3227-3238).
//  time_attn_kernel  : K1, the time-decode attention of one new frame over the growing KV cache
//      (query length 1 per (token, head); reference dreamer4.py:2021-2035, 2848, 3010).  HBM-bound: it streams
//      2*t*d*4 bytes per (token, kv head) and appends the new key/value in place on the clean pass.
#include <stdlib.h>
#include "kernels.h"
#include <float.h>

int d4_time_attn_bulk(const TimeAttnArgs& a, cudaStream_t s);   // attn_bulk.cu

namespace {

// torch.lerp(start, end, w) as ATen evaluates it (w < 0.5 ? start + w*(end-start) : end - (end-start)*(1-w))
__device__ __forceinline__ float lerp_(float a, float b, float w) {
    const float d = b - a;
    return (w < 0.5f) ? a + w * d : b - d * (1.f - w);
}

// -------------------------------------------------------------------------------------------------
// generic small attention

constexpr int SA_WARPS = 4;

__global__ void __launch_bounds__(SA_WARPS * 32) small_attn_kernel(SmallAttnArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * SA_WARPS + warp;
    if (item >= (long long)a.nb * a.hkv) return;
    const int b = (int)(item / a.hkv), hk = (int)(item % a.hkv);
    const int d = a.d, n = a.n, kp = d + 4, d4 = d >> 2;
    float* Ks = smem + (size_t)warp * 2 * n * kp;
    float* Vs = Ks + (size_t)n * kp;

    // stage K, V (value-residual lerp applied on the way in)
    for (int idx = lane; idx < n * d4; idx += 32) {
        const int j = idx / d4, c = (idx % d4) * 4;
        const float4 kv = *reinterpret_cast<const float4*>(a.k + b * a.k_sb + j * a.k_sj + (long long)hk * d + c);
        float4 vv = *reinterpret_cast<const float4*>(a.v + b * a.v_sb + j * a.v_sj + (long long)hk * d + c);
        if (a.v0) {
            const float4 rv = *reinterpret_cast<const float4*>(a.v0 + b * a.v0_sb + j * a.v0_sj + (long long)hk * d + c);
            const float w = sigmoidf_(a.mix[b * a.mix_sb + j * a.mix_sj + hk]);
            vv.x = lerp_(vv.x, rv.x, w); vv.y = lerp_(vv.y, rv.y, w); vv.z = lerp_(vv.z, rv.z, w); vv.w = lerp_(vv.w, rv.w, w);
        }
        *reinterpret_cast<float4*>(Ks + j * kp + c) = kv;
        *reinterpret_cast<float4*>(Vs + j * kp + c) = vv;
    }
    __syncwarp();
    // MultiHeadRMSNorm on keys: l2norm(k) * (gamma + 1) * sqrt(d)   (reference dreamer4.py:1663-1679)
    const float sqrt_d = sqrtf((float)d);
    for (int j = lane; j < n; j += 32) {
        float* kr = Ks + j * kp;
        float ss = 0.f;
        for (int c = 0; c < d; c += 4) { const float4 t = *reinterpret_cast<const float4*>(kr + c); ss += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w; }
        const float den = fmaxf(sqrtf(ss), D4_L2_EPS);
        for (int c = 0; c < d; ++c) kr[c] = (kr[c] / den) * ((a.k_gamma[hk * d + c] + 1.f) * sqrt_d);
    }
    __syncwarp();

    for (int gi = 0; gi < a.g; ++gi) {
        const int hq = hk * a.g + gi;
        for (int i = 0; i < a.nq; ++i) {
            const float* qp = a.q + b * a.q_sb + i * a.q_si + (long long)hq * d;
            float sc[2], p[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int j = lane + 32 * r;
                float s = -INFINITY;
                if (j < n) {
                    const float* kr = Ks + j * kp;
                    float acc = 0.f;
                    for (int c = 0; c < d; c += 4) {
                        const float4 qv = __ldg(reinterpret_cast<const float4*>(qp + c));
                        const float4 kv = *reinterpret_cast<const float4*>(kr + c);
                        acc = fmaf(qv.x, kv.x, acc); acc = fmaf(qv.y, kv.y, acc); acc = fmaf(qv.z, kv.z, acc); acc = fmaf(qv.w, kv.w, acc);
                    }
                    s = acc * a.scale;
                    if (a.softclamp > 0.f) s = tanhf(s / a.softclamp) * a.softclamp;
                    if (a.mask_agent && i < a.nq - 1 && j == n - 1) s = -FLT_MAX;
                }
                sc[r] = s;
            }
            const float mx = warp_max(fmaxf(sc[0], sc[1]));
#pragma unroll
            for (int r = 0; r < 2; ++r) p[r] = (lane + 32 * r < n) ? expf(sc[r] - mx) : 0.f;
            const float inv = 1.f / warp_sum(p[0] + p[1]);
            p[0] *= inv; p[1] *= inv;

            // out[c] = sum_j p_j V[j][c]; lane owns c = lane + 32*e
            float o[4] = {0.f, 0.f, 0.f, 0.f};
            for (int j = 0; j < n; ++j) {
                const float pj = __shfl_sync(D4_FULL, (j < 32) ? p[0] : p[1], j & 31);
#pragma unroll
                for (int e = 0; e < 4; ++e) { const int c = lane + 32 * e; if (c < d) o[e] = fmaf(pj, Vs[j * kp + c], o[e]); }
            }
            if (a.belief) {   // out -= (out . vhat) vhat,  vhat = l2norm(v_i)
                float vv[4], ss = 0.f, dot = 0.f;
#pragma unroll
                for (int e = 0; e < 4; ++e) { const int c = lane + 32 * e; vv[e] = (c < d) ? Vs[i * kp + c] : 0.f; ss += vv[e] * vv[e]; }
                const float den = fmaxf(sqrtf(warp_sum(ss)), D4_L2_EPS);
#pragma unroll
                for (int e = 0; e < 4; ++e) { vv[e] = vv[e] / den; dot += o[e] * vv[e]; }
                dot = warp_sum(dot);
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = o[e] - dot * vv[e];
            }
            float gate = 1.f;
            if (a.gate) gate = sigmoidf_(a.gate[b * a.gate_sb + i * a.gate_si + hq]);
            float* op = a.out + b * a.out_sb + i * a.out_si + (long long)hq * d;
#pragma unroll
            for (int e = 0; e < 4; ++e) { const int c = lane + 32 * e; if (c < d) op[c] = o[e] * gate; }
        }
    }
}

// -------------------------------------------------------------------------------------------------
// space attention (nq == n <= 16): one warp per (frame, kv head), every phase lane-parallel.
//   stage   K, V (value-residual lerp), Q pre-multiplied by the key-norm gain (gamma+1)*sqrt(d)   -> shared memory
//   scores  lane = (query, key) pair, 64-long dot products from shared memory; key l2-norm applied as a per-key factor
//   softmax lane = query row
//   AV      lane = 2 output columns, all queries accumulated at once from the transposed probability tile
//   belief projection, head gate, coalesced row stores
template <int D>
struct SpaceSmem {
    static constexpr int P = D + 4, SMAX = 16;
    float q[SMAX * P], k[SMAX * P], v[SMAX * P];
    float pt[SMAX * SMAX];      // probabilities, transposed: pt[j * 16 + i]
    float kinv[SMAX], vinv[SMAX];
};

constexpr int SP_WARPS = 4;

template <int D>
__global__ void __launch_bounds__(SP_WARPS * 32) space_attn_kernel(SmallAttnArgs a) {
    using SM = SpaceSmem<D>;
    constexpr int P = SM::P, C4 = D / 4, SMAX = SM::SMAX;
    constexpr int CPL = (D >= 32) ? D / 32 : 1;          // output columns per lane
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * SP_WARPS + warp;
    if (item >= (long long)a.nb * a.hkv) return;
    const int b = (int)(item / a.hkv), hk = (int)(item % a.hkv);
    SM& sm = reinterpret_cast<SM*>(smem_raw)[warp];
    const int S = a.n;
    const float sqrt_d = sqrtf((float)D);

    for (int idx = lane; idx < SMAX * SMAX; idx += 32) sm.pt[idx] = 0.f;
    if (a.v0 && lane < S) sm.kinv[lane] = sigmoidf_(a.mix[b * a.mix_sb + lane * a.mix_sj + hk]);   // value-residual mix weight per key (kinv is free until the norms)
    __syncwarp();
    for (int idx = lane; idx < S * C4; idx += 32) {
        const int j = idx / C4, c = (idx % C4) * 4;
        const float4 kv = *reinterpret_cast<const float4*>(a.k + b * a.k_sb + j * a.k_sj + (long long)hk * D + c);
        float4 vv = *reinterpret_cast<const float4*>(a.v + b * a.v_sb + j * a.v_sj + (long long)hk * D + c);
        if (a.v0) {
            const float4 rv = *reinterpret_cast<const float4*>(a.v0 + b * a.v0_sb + j * a.v0_sj + (long long)hk * D + c);
            const float w = sm.kinv[j];
            vv.x = lerp_(vv.x, rv.x, w); vv.y = lerp_(vv.y, rv.y, w); vv.z = lerp_(vv.z, rv.z, w); vv.w = lerp_(vv.w, rv.w, w);
        }
        *reinterpret_cast<float4*>(sm.k + j * P + c) = kv;
        *reinterpret_cast<float4*>(sm.v + j * P + c) = vv;
    }
    __syncwarp();
    {   // l2 norms: lanes 0..15 keys, lanes 16..31 values
        const int j = lane & 15;
        if (j < S) {
            const float* r = (lane < 16 ? sm.k : sm.v) + j * P;
            float ss = 0.f;
#pragma unroll
            for (int c = 0; c < D; c += 4) { const float4 t = *reinterpret_cast<const float4*>(r + c); ss += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w; }
            const float inv = 1.f / fmaxf(sqrtf(ss), D4_L2_EPS);
            (lane < 16 ? sm.kinv : sm.vinv)[j] = inv;
        }
    }

    for (int gi = 0; gi < a.g; ++gi) {
        const int hq = hk * a.g + gi;
        __syncwarp();
        for (int idx = lane; idx < S * C4; idx += 32) {
            const int i = idx / C4, c = (idx % C4) * 4;
            float4 qv = *reinterpret_cast<const float4*>(a.q + b * a.q_sb + i * a.q_si + (long long)hq * D + c);
            const float4 gm = *reinterpret_cast<const float4*>(a.k_gamma + hk * D + c);
            qv.x *= (gm.x + 1.f) * sqrt_d; qv.y *= (gm.y + 1.f) * sqrt_d; qv.z *= (gm.z + 1.f) * sqrt_d; qv.w *= (gm.w + 1.f) * sqrt_d;
            *reinterpret_cast<float4*>(sm.q + i * P + c) = qv;
        }
        __syncwarp();
        for (int pair = lane; pair < S * S; pair += 32) {
            const int i = pair / S, j = pair - i * S;
            const float* qr = sm.q + i * P; const float* kr = sm.k + j * P;
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < D; c += 4) {
                const float4 qv = *reinterpret_cast<const float4*>(qr + c);
                const float4 kv = *reinterpret_cast<const float4*>(kr + c);
                acc = fmaf(qv.x, kv.x, acc); acc = fmaf(qv.y, kv.y, acc); acc = fmaf(qv.z, kv.z, acc); acc = fmaf(qv.w, kv.w, acc);
            }
            float s = acc * sm.kinv[j] * a.scale;
            if (a.softclamp > 0.f) s = tanhf(s / a.softclamp) * a.softclamp;
            if (a.mask_agent && i < S - 1 && j == S - 1) s = -FLT_MAX;
            sm.pt[j * SMAX + i] = s;
        }
        __syncwarp();
        if (lane < S) {
            float mx = -INFINITY;
            for (int j = 0; j < S; ++j) mx = fmaxf(mx, sm.pt[j * SMAX + lane]);
            float sum = 0.f;
            for (int j = 0; j < S; ++j) { const float e = expf(sm.pt[j * SMAX + lane] - mx); sm.pt[j * SMAX + lane] = e; sum += e; }
            const float inv = 1.f / sum;
            for (int j = 0; j < S; ++j) sm.pt[j * SMAX + lane] *= inv;
        }
        __syncwarp();
        float acc[SMAX][CPL];
#pragma unroll
        for (int i = 0; i < SMAX; ++i)
#pragma unroll
            for (int e = 0; e < CPL; ++e) acc[i][e] = 0.f;
        const int c0 = lane * CPL;
        const bool col_ok = c0 < D;
        for (int j = 0; j < S; ++j) {
            float vv[CPL];
#pragma unroll
            for (int e = 0; e < CPL; ++e) vv[e] = col_ok ? sm.v[j * P + c0 + e] : 0.f;
#pragma unroll
            for (int i4 = 0; i4 < SMAX; i4 += 4) {
                const float4 p4 = *reinterpret_cast<const float4*>(sm.pt + j * SMAX + i4);
#pragma unroll
                for (int e = 0; e < CPL; ++e) {
                    acc[i4 + 0][e] = fmaf(p4.x, vv[e], acc[i4 + 0][e]); acc[i4 + 1][e] = fmaf(p4.y, vv[e], acc[i4 + 1][e]);
                    acc[i4 + 2][e] = fmaf(p4.z, vv[e], acc[i4 + 2][e]); acc[i4 + 3][e] = fmaf(p4.w, vv[e], acc[i4 + 3][e]);
                }
            }
        }
        const float gate_l = (a.gate && lane < S) ? sigmoidf_(a.gate[b * a.gate_sb + lane * a.gate_si + hq]) : 1.f;
#pragma unroll
        for (int i = 0; i < SMAX; ++i) {
            if (i < S) {                                   // warp-uniform
                float o[CPL];
#pragma unroll
                for (int e = 0; e < CPL; ++e) o[e] = acc[i][e];
                if (a.belief) {   // out -= (out . vhat) vhat,  vhat = l2norm(v_i)
                    float vh[CPL], dot = 0.f;
                    const float vinv = sm.vinv[i];
#pragma unroll
                    for (int e = 0; e < CPL; ++e) { vh[e] = col_ok ? sm.v[i * P + c0 + e] * vinv : 0.f; dot = fmaf(o[e], vh[e], dot); }
                    dot = warp_sum(dot);
#pragma unroll
                    for (int e = 0; e < CPL; ++e) o[e] = o[e] - dot * vh[e];
                }
                const float gate = __shfl_sync(D4_FULL, gate_l, i);
                if (col_ok) {
                    float* op = a.out + b * a.out_sb + i * a.out_si + (long long)hq * D + c0;
#pragma unroll
                    for (int e = 0; e < CPL; ++e) op[e] = o[e] * gate;
                }
            }
        }
    }
}

// -------------------------------------------------------------------------------------------------
// space attention on the warp-level tensor cores (tf32x3 / tf32 engine modes): the same S x S (S <= 16) attention as
// space_attn_kernel, with Q K^T and P V as m16n8k8 TF32 mma.sync tiles and the 3-term TF32 split (operands rounded to
// nearest, fp32 accumulate) that keeps them fp32-accurate.  One warp per (frame, kv head): a 16 x 16 x D score tile and a
// 16 x D x 16 output tile are far below a tcgen05 tile (M = 128 rows of ONE operand pair), so the warp-level MMA is the
// instruction that fits; it cuts the kernel from ~4600 to ~1600 issued instructions per (frame, head).
//   fragment layouts (PTX ISA, m16n8k8 .tf32): g = lane / 4, t = lane % 4
//     A (16 x 8):  a0 (g, t)  a1 (g + 8, t)  a2 (g, t + 4)  a3 (g + 8, t + 4)
//     B ( 8 x 8):  b0 (k = t, n = g)  b1 (k = t + 4, n = g)
//     C (16 x 8):  c0 (g, 2t)  c1 (g, 2t + 1)  c2 (g + 8, 2t)  c3 (g + 8, 2t + 1)
// ALIAS (one query group): V is fetched into registers up front and written over K's tile once the scores are done,
// which cuts the per-warp shared memory from 14.8 to 10.5 KB (5 resident CTAs instead of 3) and lets Q, K, V and the value
// residual all be in flight behind ONE exposed global-memory latency.
template <int D, bool ALIAS>
struct SpaceMmaSmem {
    static constexpr int PQ = D + 4, PV = D + 8, PP = 20;
    float q[16 * PQ];
    float k[16 * (ALIAS ? PV : PQ)];             // keys (pitch PQ); with ALIAS the values (pitch PV) replace them after the scores
    float v[ALIAS ? 4 : 16 * PV];
    float p[16 * PP];
    float kinv[16], vinv[16], gate[16], mixw[16];
};

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    const float h = tf32_rna(x);
    hi = __float_as_uint(h);
    lo = __float_as_uint(tf32_rna(x - h));
}
// c += a * b with a, b given as fp32 fragments: a_lo*b_hi + a_hi*b_lo + a_hi*b_hi
__device__ __forceinline__ void mma_3x(float (&c)[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4], const uint32_t (&bhi)[2], const uint32_t (&blo)[2]) {
    mma_tf32(c, alo, bhi);
    mma_tf32(c, ahi, blo);
    mma_tf32(c, ahi, bhi);
}

template <int D, bool ALIAS>
__global__ void __launch_bounds__(SP_WARPS * 32, ALIAS ? 4 : 3) space_attn_mma_kernel(SmallAttnArgs a) {
    using SM = SpaceMmaSmem<D, ALIAS>;
    constexpr int PQ = SM::PQ, PV = SM::PV, PP = SM::PP, C4 = D / 4, KS = D / 8, NT = D / 8;
    constexpr int VR = (16 * C4 + 31) / 32;          // float4 values per lane when V is held in registers
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * SP_WARPS + warp;
    if (item >= (long long)a.nb * a.hkv) return;
    const int b = (int)(item / a.hkv), hk = (int)(item % a.hkv);
    SM& sm = reinterpret_cast<SM*>(smem_raw)[warp];
    float* smv = ALIAS ? sm.k : sm.v;                // where the PV pass finds the values
    const int S = a.n;
    const int g = lane >> 2, t = lane & 3;
    const float sqrt_d = sqrtf((float)D);

    auto stage_q = [&](int hq) {      // Q pre-multiplied by the key-norm gain (gamma + 1) * sqrt(d); rows S..15 zero; head gates
        for (int idx = lane; idx < 16 * C4; idx += 32) {
            const int i = idx / C4, c = (idx % C4) * 4;
            float4 qv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < S) {
                qv = *reinterpret_cast<const float4*>(a.q + b * a.q_sb + i * a.q_si + (long long)hq * D + c);
                const float4 gm = *reinterpret_cast<const float4*>(a.k_gamma + hk * D + c);
                qv.x *= (gm.x + 1.f) * sqrt_d; qv.y *= (gm.y + 1.f) * sqrt_d; qv.z *= (gm.z + 1.f) * sqrt_d; qv.w *= (gm.w + 1.f) * sqrt_d;
            }
            *reinterpret_cast<float4*>(sm.q + i * PQ + c) = qv;
        }
        if (lane < 16) sm.gate[lane] = (a.gate && lane < S) ? sigmoidf_(a.gate[b * a.gate_sb + lane * a.gate_si + hq]) : 1.f;
    };

    // ---- everything this (frame, head) needs is requested before the first wait: mix logits, K, V (+ value residual), Q
    const float mixl = (a.v0 && lane < S) ? a.mix[b * a.mix_sb + lane * a.mix_sj + hk] : 0.f;
    float4 vreg[ALIAS ? VR : 1], rreg[ALIAS ? VR : 1];
#pragma unroll
    for (int r = 0; r < VR; ++r) {
        const int idx = lane + 32 * r, j = idx / C4, c = (idx % C4) * 4;
        float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv, rv = kv;
        if (idx < 16 * C4 && j < S) {
            kv = *reinterpret_cast<const float4*>(a.k + b * a.k_sb + j * a.k_sj + (long long)hk * D + c);
            vv = *reinterpret_cast<const float4*>(a.v + b * a.v_sb + j * a.v_sj + (long long)hk * D + c);
            if (a.v0) rv = *reinterpret_cast<const float4*>(a.v0 + b * a.v0_sb + j * a.v0_sj + (long long)hk * D + c);
        }
        if (idx < 16 * C4) *reinterpret_cast<float4*>(sm.k + j * PQ + c) = kv;
        if (ALIAS) { vreg[r] = vv; rreg[r] = rv; }
        else if (idx < 16 * C4) { *reinterpret_cast<float4*>(sm.v + j * PV + c) = vv; *reinterpret_cast<float4*>(sm.q + j * PQ + c) = rv; }   // residual parked in q's tile
    }
    if (lane < 16) sm.mixw[lane] = sigmoidf_(mixl);
    if (ALIAS) stage_q(hk);
    __syncwarp();
    // value-residual lerp (values still in registers / in their own tile)
    if (a.v0) {
#pragma unroll
        for (int r = 0; r < VR; ++r) {
            const int idx = lane + 32 * r, j = idx / C4, c = (idx % C4) * 4;
            if (idx < 16 * C4 && j < S) {
                const float w = sm.mixw[j];
                if (ALIAS) {
                    vreg[r].x = lerp_(vreg[r].x, rreg[r].x, w); vreg[r].y = lerp_(vreg[r].y, rreg[r].y, w);
                    vreg[r].z = lerp_(vreg[r].z, rreg[r].z, w); vreg[r].w = lerp_(vreg[r].w, rreg[r].w, w);
                } else {
                    float4 vv = *reinterpret_cast<float4*>(sm.v + j * PV + c);
                    const float4 rv = *reinterpret_cast<const float4*>(sm.q + j * PQ + c);
                    vv.x = lerp_(vv.x, rv.x, w); vv.y = lerp_(vv.y, rv.y, w); vv.z = lerp_(vv.z, rv.z, w); vv.w = lerp_(vv.w, rv.w, w);
                    *reinterpret_cast<float4*>(sm.v + j * PV + c) = vv;
                }
            }
        }
    }
    if (lane < 16) {   // key l2 norms
        const float* r = sm.k + lane * PQ;
        float ss = 0.f;
#pragma unroll
        for (int c = 0; c < D; c += 4) { const float4 x = *reinterpret_cast<const float4*>(r + c); ss += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w; }
        sm.kinv[lane] = 1.f / fmaxf(sqrtf(ss), D4_L2_EPS);
    }
    if (!ALIAS) {
        __syncwarp();
        if (lane < 16) {
            const float* r = sm.v + lane * PV;
            float ss = 0.f;
#pragma unroll
            for (int c = 0; c < D; c += 4) { const float4 x = *reinterpret_cast<const float4*>(r + c); ss += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w; }
            sm.vinv[lane] = 1.f / fmaxf(sqrtf(ss), D4_L2_EPS);
        }
    }

    for (int gi = 0; gi < a.g; ++gi) {
        const int hq = hk * a.g + gi;
        __syncwarp();
        if (!ALIAS) stage_q(hq);
        __syncwarp();

        // ---- scores: 16 queries x 16 keys, two 8-key tiles
        float sc[2][4];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int r = 0; r < 4; ++r) sc[nt][r] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            uint32_t ahi[4], alo[4];
            split_tf32(sm.q[g * PQ + ks * 8 + t], ahi[0], alo[0]);
            split_tf32(sm.q[(g + 8) * PQ + ks * 8 + t], ahi[1], alo[1]);
            split_tf32(sm.q[g * PQ + ks * 8 + t + 4], ahi[2], alo[2]);
            split_tf32(sm.q[(g + 8) * PQ + ks * 8 + t + 4], ahi[3], alo[3]);
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                uint32_t bhi[2], blo[2];
                split_tf32(sm.k[(nt * 8 + g) * PQ + ks * 8 + t], bhi[0], blo[0]);
                split_tf32(sm.k[(nt * 8 + g) * PQ + ks * 8 + t + 4], bhi[1], blo[1]);
                mma_3x(sc[nt], ahi, alo, bhi, blo);
            }
        }
        // scale by the key norm, softclamp, mask, softmax over the keys of each query row (rows g and g + 8)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int i = g + ((r & 2) ? 8 : 0), j = nt * 8 + 2 * t + (r & 1);
                float s = sc[nt][r] * sm.kinv[j] * a.scale;
                if (a.softclamp > 0.f) s = tanhf(s / a.softclamp) * a.softclamp;
                if (a.mask_agent && i < S - 1 && j == S - 1) s = -FLT_MAX;
                if (j >= S) s = -INFINITY;
                sc[nt][r] = s;
            }
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {          // half 0: row g (regs 0, 1), half 1: row g + 8 (regs 2, 3)
            float mx = fmaxf(fmaxf(sc[0][2 * half], sc[0][2 * half + 1]), fmaxf(sc[1][2 * half], sc[1][2 * half + 1]));
            mx = fmaxf(mx, __shfl_xor_sync(D4_FULL, mx, 1)); mx = fmaxf(mx, __shfl_xor_sync(D4_FULL, mx, 2));
            float e[4], sum = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) { e[q] = expf(sc[q >> 1][2 * half + (q & 1)] - mx); sum += e[q]; }
            sum += __shfl_xor_sync(D4_FULL, sum, 1); sum += __shfl_xor_sync(D4_FULL, sum, 2);
            const float inv = 1.f / sum;
            const int i = g + 8 * half;
#pragma unroll
            for (int q = 0; q < 4; ++q) sm.p[i * PP + (q >> 1) * 8 + 2 * t + (q & 1)] = e[q] * inv;
        }
        __syncwarp();
        if (ALIAS) {   // the keys are dead: the (lerped) values take their tile, and their row norms come from the registers
#pragma unroll
            for (int r = 0; r < VR; ++r) {
                const int idx = lane + 32 * r, j = idx / C4, c = (idx % C4) * 4;
                if (idx < 16 * C4) {
                    *reinterpret_cast<float4*>(smv + j * PV + c) = vreg[r];
                    float ss = vreg[r].x * vreg[r].x + vreg[r].y * vreg[r].y + vreg[r].z * vreg[r].z + vreg[r].w * vreg[r].w;
#pragma unroll
                    for (int off = C4 / 2; off > 0; off >>= 1) ss += __shfl_xor_sync(D4_FULL, ss, off);     // the C4 lanes that share row j
                    if ((lane & (C4 - 1)) == 0) sm.vinv[j] = 1.f / fmaxf(sqrtf(ss), D4_L2_EPS);
                }
            }
            __syncwarp();
        }

        // ---- out = P V : 16 queries x D, K = 16 keys (two k-steps), D / 8 column tiles
        uint32_t phi[2][4], plo[2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            split_tf32(sm.p[g * PP + ks * 8 + t], phi[ks][0], plo[ks][0]);
            split_tf32(sm.p[(g + 8) * PP + ks * 8 + t], phi[ks][1], plo[ks][1]);
            split_tf32(sm.p[g * PP + ks * 8 + t + 4], phi[ks][2], plo[ks][2]);
            split_tf32(sm.p[(g + 8) * PP + ks * 8 + t + 4], phi[ks][3], plo[ks][3]);
        }
        float o[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
            for (int r = 0; r < 4; ++r) o[nt][r] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                uint32_t bhi[2], blo[2];
                split_tf32(smv[(ks * 8 + t) * PV + nt * 8 + g], bhi[0], blo[0]);
                split_tf32(smv[(ks * 8 + t + 4) * PV + nt * 8 + g], bhi[1], blo[1]);
                mma_3x(o[nt], phi[ks], plo[ks], bhi, blo);
            }
        }
        // ---- belief projection (out -= (out . vhat) vhat, vhat = l2norm(v_i)), head gate, store rows g and g + 8
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int i = g + 8 * half;
            const float* vr = smv + i * PV;
            float dot = 0.f;
            if (a.belief) {
                const float vinv = sm.vinv[i];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const float2 vv = *reinterpret_cast<const float2*>(vr + nt * 8 + 2 * t);
                    dot = fmaf(o[nt][2 * half], vv.x * vinv, dot); dot = fmaf(o[nt][2 * half + 1], vv.y * vinv, dot);
                }
                dot += __shfl_xor_sync(D4_FULL, dot, 1); dot += __shfl_xor_sync(D4_FULL, dot, 2);
            }
            if (i < S) {
                const float gate = sm.gate[i];
                const float vinv = a.belief ? sm.vinv[i] : 0.f;
                float* op = a.out + b * a.out_sb + i * a.out_si + (long long)hq * D;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const float2 vv = *reinterpret_cast<const float2*>(vr + nt * 8 + 2 * t);
                    float2 r;
                    r.x = (o[nt][2 * half] - dot * (vv.x * vinv)) * gate;
                    r.y = (o[nt][2 * half + 1] - dot * (vv.y * vinv)) * gate;
                    *reinterpret_cast<float2*>(op + nt * 8 + 2 * t) = r;
                }
            }
        }
    }
}

// -------------------------------------------------------------------------------------------------
// attention-residual pool (one query per token, 4 heads x 64 = 256 wide, n <= 32 context hiddens): one warp per token
// does all heads at once straight from global memory — lane = (head = lane / 8, 8-float slice of the head's 64 dims),
// key l2-norm and q.k reduced over the 8 lanes of a head with shuffles, single-pass online softmax.
constexpr int PL_WARPS = 8;

__global__ void __launch_bounds__(PL_WARPS * 32) pool_attn_kernel(SmallAttnArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long tok = (long long)blockIdx.x * PL_WARPS + warp;
    if (tok >= a.nb) return;
    const int c = lane * 8;                         // column slice [c, c+8) of the 256-wide row; head = lane / 8
    const float sqrt_d = 8.f;
    float q[8];
    {
        const float4 q0 = *reinterpret_cast<const float4*>(a.q + tok * a.q_sb + c), q1 = *reinterpret_cast<const float4*>(a.q + tok * a.q_sb + c + 4);
        const float4 g0 = *reinterpret_cast<const float4*>(a.k_gamma + c), g1 = *reinterpret_cast<const float4*>(a.k_gamma + c + 4);
        q[0] = q0.x * ((g0.x + 1.f) * sqrt_d); q[1] = q0.y * ((g0.y + 1.f) * sqrt_d); q[2] = q0.z * ((g0.z + 1.f) * sqrt_d); q[3] = q0.w * ((g0.w + 1.f) * sqrt_d);
        q[4] = q1.x * ((g1.x + 1.f) * sqrt_d); q[5] = q1.y * ((g1.y + 1.f) * sqrt_d); q[6] = q1.z * ((g1.z + 1.f) * sqrt_d); q[7] = q1.w * ((g1.w + 1.f) * sqrt_d);
    }
    float mx = -INFINITY, den = 0.f, o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = 0.f;
    const float* kp = a.k + tok * a.k_sb + c;
    const float* vp = a.v + tok * a.v_sb + c;
#pragma unroll 2
    for (int j = 0; j < a.n; ++j) {
        const float4 k0 = __ldcs(reinterpret_cast<const float4*>(kp + j * a.k_sj)), k1 = __ldcs(reinterpret_cast<const float4*>(kp + j * a.k_sj + 4));
        const float4 v0 = __ldcs(reinterpret_cast<const float4*>(vp + j * a.v_sj)), v1 = __ldcs(reinterpret_cast<const float4*>(vp + j * a.v_sj + 4));
        float ss = k0.x * k0.x + k0.y * k0.y + k0.z * k0.z + k0.w * k0.w + k1.x * k1.x + k1.y * k1.y + k1.z * k1.z + k1.w * k1.w;
        float dot = q[0] * k0.x + q[1] * k0.y + q[2] * k0.z + q[3] * k0.w + q[4] * k1.x + q[5] * k1.y + q[6] * k1.z + q[7] * k1.w;
#pragma unroll
        for (int off = 4; off > 0; off >>= 1) { ss += __shfl_xor_sync(D4_FULL, ss, off); dot += __shfl_xor_sync(D4_FULL, dot, off); }
        const float s = dot / fmaxf(sqrtf(ss), D4_L2_EPS) * a.scale;
        const float nmx = fmaxf(mx, s);
        const float corr = expf(mx - nmx), p = expf(s - nmx);       // mx = -inf on the first key: corr = 0
        den = den * corr + p;
        o[0] = fmaf(p, v0.x, o[0] * corr); o[1] = fmaf(p, v0.y, o[1] * corr); o[2] = fmaf(p, v0.z, o[2] * corr); o[3] = fmaf(p, v0.w, o[3] * corr);
        o[4] = fmaf(p, v1.x, o[4] * corr); o[5] = fmaf(p, v1.y, o[5] * corr); o[6] = fmaf(p, v1.z, o[6] * corr); o[7] = fmaf(p, v1.w, o[7] * corr);
        mx = nmx;
    }
    float gate = 1.f;
    if (a.gate_w) {          // gate logits of the 4 heads from the (RMS-normalised) token itself
        const float* xr = a.gate_x + tok * a.gate_x_ld;
        float g4[4] = {0.f, 0.f, 0.f, 0.f};
        for (int cc = lane * 4; cc < a.gate_D; cc += 128) {
            const float4 xv = *reinterpret_cast<const float4*>(xr + cc);
#pragma unroll
            for (int hh = 0; hh < 4; ++hh) {
                const float4 wv = __ldg(reinterpret_cast<const float4*>(a.gate_w + (long long)hh * a.gate_D + cc));
                g4[hh] = fmaf(xv.x, wv.x, g4[hh]); g4[hh] = fmaf(xv.y, wv.y, g4[hh]); g4[hh] = fmaf(xv.z, wv.z, g4[hh]); g4[hh] = fmaf(xv.w, wv.w, g4[hh]);
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
            for (int hh = 0; hh < 4; ++hh) g4[hh] += __shfl_xor_sync(D4_FULL, g4[hh], off);
        }
        const int hh = lane >> 3;
        float rstd = a.gate_rstd[tok];
        if (a.gate_rstd_is_ss) rstd = rsqrtf(rstd / (float)a.gate_D + D4_RMS_EPS);
        const float logit = (hh == 0 ? g4[0] : hh == 1 ? g4[1] : hh == 2 ? g4[2] : g4[3]) * rstd;
        gate = sigmoidf_(logit);
    } else if (a.gate) {
        gate = sigmoidf_(a.gate[tok * a.gate_sb + (lane >> 3)]);
    }
    const float sc = gate / den;
    float* op = a.out + tok * a.out_sb + c;
    *reinterpret_cast<float4*>(op) = make_float4(o[0] * sc, o[1] * sc, o[2] * sc, o[3] * sc);
    *reinterpret_cast<float4*>(op + 4) = make_float4(o[4] * sc, o[5] * sc, o[6] * sc, o[7] * sc);
}

// -------------------------------------------------------------------------------------------------
// K1: time-decode attention.  One warp per (token, kv head) stream.  HALF = d/2 rotary pairs; lane p owns
// elements (p, p + HALF) of every d-vector (the rotate-half partner lives in the same lane).

template <int D, int G>
struct TimeAttnSmem {
    float tile[32 * D];     // one staged tile of 32 keys (or values), linear [key][d]
    float q[G * D];         // rotated queries of the group
};

constexpr int TA_WARPS = 8;
constexpr int TA_MAXTILES = 8;       // up to 256 cached frames

template <int D, int G>
__global__ void __launch_bounds__(TA_WARPS * 32) time_attn_kernel(TimeAttnArgs a) {
    constexpr int HALF = D / 2;
    constexpr int PPL = (HALF + 31) / 32;          // rotary pairs per lane
    constexpr int C4 = D / 4;                      // float4 chunks per key
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * TA_WARPS + warp;
    if (item >= (long long)a.M * a.hkv) return;
    const int m = (int)(item / a.hkv), hk = (int)(item % a.hkv);
    TimeAttnSmem<D, G>& sm = reinterpret_cast<TimeAttnSmem<D, G>*>(smem_raw)[warp];

    const float* row = a.qkvgm + (long long)m * a.ld;
    const float sqrt_d = sqrtf((float)D);

    // ---- prologue: new k, v (lerp with value residual), key head-norm, rotary on q and k
    float k1[PPL], k2[PPL], v1[PPL], v2[PPL], cs[PPL], sn[PPL];
    const float mixw = sigmoidf_(row[a.off_m + hk]);
    float ss = 0.f;
#pragma unroll
    for (int e = 0; e < PPL; ++e) {
        const int p = lane + 32 * e;
        k1[e] = k2[e] = v1[e] = v2[e] = 0.f; cs[e] = 1.f; sn[e] = 0.f;
        if (p < HALF) {
            k1[e] = row[a.off_k + hk * D + p]; k2[e] = row[a.off_k + hk * D + p + HALF];
            const float r1 = a.v0[(long long)m * a.ldv0 + hk * D + p], r2 = a.v0[(long long)m * a.ldv0 + hk * D + p + HALF];
            v1[e] = lerp_(row[a.off_v + hk * D + p], r1, mixw);
            v2[e] = lerp_(row[a.off_v + hk * D + p + HALF], r2, mixw);
            ss += k1[e] * k1[e] + k2[e] * k2[e];
            const float ang = (float)a.t * a.inv_freq[p];
            sincosf(ang, &sn[e], &cs[e]);
        }
    }
    const float kden = fmaxf(sqrtf(warp_sum(ss)), D4_L2_EPS);
#pragma unroll
    for (int e = 0; e < PPL; ++e) {
        const int p = lane + 32 * e;
        if (p < HALF) {
            const float n1 = (k1[e] / kden) * ((a.k_gamma[hk * D + p] + 1.f) * sqrt_d);
            const float n2 = (k2[e] / kden) * ((a.k_gamma[hk * D + p + HALF] + 1.f) * sqrt_d);
            k1[e] = n1 * cs[e] + (-n2) * sn[e];
            k2[e] = n2 * cs[e] + n1 * sn[e];
        }
    }
    float self_s[G];
#pragma unroll
    for (int gi = 0; gi < G; ++gi) {
        const int hq = hk * G + gi;
        float dot = 0.f;
#pragma unroll
        for (int e = 0; e < PPL; ++e) {
            const int p = lane + 32 * e;
            if (p < HALF) {
                const float q1 = row[hq * D + p], q2 = row[hq * D + p + HALF];
                const float r1 = q1 * cs[e] + (-q2) * sn[e];
                const float r2 = q2 * cs[e] + q1 * sn[e];
                sm.q[gi * D + p] = r1; sm.q[gi * D + p + HALF] = r2;
                dot += r1 * k1[e] + r2 * k2[e];
            }
        }
        float s = warp_sum(dot) * a.scale;
        if (a.softclamp > 0.f) s = tanhf(s / a.softclamp) * a.softclamp;
        self_s[gi] = s;
    }
    __syncwarp();

    // ---- scores over the cached keys: coalesced 128-bit loads -> smem tile -> lane-per-key dot products
    const int t = a.t;
    const int ntiles = (t + 31) / 32;
    const float* kc = a.kcache + ((long long)m * a.hkv + hk) * (long long)a.Tmax * D;
    const float* vc = a.vcache + ((long long)m * a.hkv + hk) * (long long)a.Tmax * D;
    float sc[G][TA_MAXTILES];
#pragma unroll
    for (int ti = 0; ti < TA_MAXTILES; ++ti) {
#pragma unroll
        for (int gi = 0; gi < G; ++gi) sc[gi][ti] = -INFINITY;
        if (ti < ntiles) {
            const int nk = min(32, t - ti * 32);
            const float4* src = reinterpret_cast<const float4*>(kc + (long long)ti * 32 * D);
            float4* dst = reinterpret_cast<float4*>(sm.tile);
            for (int idx = lane; idx < nk * C4; idx += 32) dst[idx] = __ldcs(src + idx);
            __syncwarp();
            if (lane < nk) {
                float acc[G];
#pragma unroll
                for (int gi = 0; gi < G; ++gi) acc[gi] = 0.f;
                // rotated chunk order keeps the linear [key][d] tile bank-conflict free for 128-bit reads
#pragma unroll
                for (int c = 0; c < C4; ++c) {
                    const int cc = (c + lane) % C4;
                    const float4 kv = *reinterpret_cast<const float4*>(sm.tile + lane * D + cc * 4);
#pragma unroll
                    for (int gi = 0; gi < G; ++gi) {
                        const float4 qv = *reinterpret_cast<const float4*>(sm.q + gi * D + cc * 4);
                        acc[gi] = fmaf(qv.x, kv.x, acc[gi]); acc[gi] = fmaf(qv.y, kv.y, acc[gi]);
                        acc[gi] = fmaf(qv.z, kv.z, acc[gi]); acc[gi] = fmaf(qv.w, kv.w, acc[gi]);
                    }
                }
#pragma unroll
                for (int gi = 0; gi < G; ++gi) {
                    float s = acc[gi] * a.scale;
                    if (a.softclamp > 0.f) s = tanhf(s / a.softclamp) * a.softclamp;
                    sc[gi][ti] = s;
                }
            }
            __syncwarp();
        }
    }

    // ---- softmax over cached keys + self
    float pself[G];
#pragma unroll
    for (int gi = 0; gi < G; ++gi) {
        float mx = self_s[gi];
#pragma unroll
        for (int ti = 0; ti < TA_MAXTILES; ++ti) mx = fmaxf(mx, sc[gi][ti]);
        mx = warp_max(mx);
        float sum = 0.f;
#pragma unroll
        for (int ti = 0; ti < TA_MAXTILES; ++ti) { const float e = (sc[gi][ti] == -INFINITY) ? 0.f : expf(sc[gi][ti] - mx); sc[gi][ti] = e; sum += e; }
        sum = warp_sum(sum);
        const float es = expf(self_s[gi] - mx);
        const float inv = 1.f / (sum + es);
#pragma unroll
        for (int ti = 0; ti < TA_MAXTILES; ++ti) sc[gi][ti] *= inv;
        pself[gi] = es * inv;
    }

    // ---- AV: stream the cached values tile by tile
    float o1[G][PPL], o2[G][PPL];
#pragma unroll
    for (int gi = 0; gi < G; ++gi)
#pragma unroll
        for (int e = 0; e < PPL; ++e) { o1[gi][e] = 0.f; o2[gi][e] = 0.f; }
#pragma unroll
    for (int ti = 0; ti < TA_MAXTILES; ++ti) {
        if (ti < ntiles) {
            const int nk = min(32, t - ti * 32);
            const float4* src = reinterpret_cast<const float4*>(vc + (long long)ti * 32 * D);
            float4* dst = reinterpret_cast<float4*>(sm.tile);
            for (int idx = lane; idx < nk * C4; idx += 32) dst[idx] = __ldcs(src + idx);
            __syncwarp();
            for (int j = 0; j < nk; ++j) {
#pragma unroll
                for (int gi = 0; gi < G; ++gi) {
                    const float pj = __shfl_sync(D4_FULL, sc[gi][ti], j);
#pragma unroll
                    for (int e = 0; e < PPL; ++e) {
                        const int p = lane + 32 * e;
                        if (p < HALF) {
                            o1[gi][e] = fmaf(pj, sm.tile[j * D + p], o1[gi][e]);
                            o2[gi][e] = fmaf(pj, sm.tile[j * D + p + HALF], o2[gi][e]);
                        }
                    }
                }
            }
            __syncwarp();
        }
    }

    // ---- epilogue: + self, belief projection on the new value, head gate, store; append k/v on the clean pass
    float vs = 0.f;
#pragma unroll
    for (int e = 0; e < PPL; ++e) vs += v1[e] * v1[e] + v2[e] * v2[e];
    const float vden = fmaxf(sqrtf(warp_sum(vs)), D4_L2_EPS);
#pragma unroll
    for (int gi = 0; gi < G; ++gi) {
        const int hq = hk * G + gi;
        float dot = 0.f;
#pragma unroll
        for (int e = 0; e < PPL; ++e) {
            o1[gi][e] = fmaf(pself[gi], v1[e], o1[gi][e]);
            o2[gi][e] = fmaf(pself[gi], v2[e], o2[gi][e]);
            dot += o1[gi][e] * (v1[e] / vden) + o2[gi][e] * (v2[e] / vden);
        }
        dot = warp_sum(dot);
        const float gate = sigmoidf_(row[a.off_g + hq]);
        float* op = a.out + (long long)m * a.ldo + hq * D;
#pragma unroll
        for (int e = 0; e < PPL; ++e) {
            const int p = lane + 32 * e;
            if (p < HALF) {
                op[p] = (o1[gi][e] - dot * (v1[e] / vden)) * gate;
                op[p + HALF] = (o2[gi][e] - dot * (v2[e] / vden)) * gate;
            }
        }
    }
    if (a.commit) {
        float* kd = a.kcache + (((long long)m * a.hkv + hk) * a.Tmax + t) * D;
        float* vd = a.vcache + (((long long)m * a.hkv + hk) * a.Tmax + t) * D;
#pragma unroll
        for (int e = 0; e < PPL; ++e) {
            const int p = lane + 32 * e;
            if (p < HALF) { kd[p] = k1[e]; kd[p + HALF] = k2[e]; vd[p] = v1[e]; vd[p + HALF] = v2[e]; }
        }
    }
}

template <int D, int G>
int launch_time_attn(const TimeAttnArgs& a, cudaStream_t s) {
    const size_t smem = sizeof(TimeAttnSmem<D, G>) * TA_WARPS;
    static bool configured = false;
    if (!configured) {
        D4_CUDA_OK(cudaFuncSetAttribute(time_attn_kernel<D, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const long long items = (long long)a.M * a.hkv;
    time_attn_kernel<D, G><<<(unsigned)((items + TA_WARPS - 1) / TA_WARPS), TA_WARPS * 32, smem, s>>>(a);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError());
    return 0;
}


template <int D>
int launch_space(const SmallAttnArgs& a, cudaStream_t s) {
    const size_t smem = sizeof(SpaceSmem<D>) * SP_WARPS;
    static bool configured = false;
    if (!configured) {
        D4_CUDA_OK(cudaFuncSetAttribute(space_attn_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const long long items = (long long)a.nb * a.hkv;
    space_attn_kernel<D><<<(unsigned)((items + SP_WARPS - 1) / SP_WARPS), SP_WARPS * 32, smem, s>>>(a);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int D, bool ALIAS>
int launch_space_mma_impl(const SmallAttnArgs& a, cudaStream_t s) {
    const size_t smem = sizeof(SpaceMmaSmem<D, ALIAS>) * SP_WARPS;
    static bool configured = false;
    if (!configured) {
        D4_CUDA_OK(cudaFuncSetAttribute(space_attn_mma_kernel<D, ALIAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const long long items = (long long)a.nb * a.hkv;
    space_attn_mma_kernel<D, ALIAS><<<(unsigned)((items + SP_WARPS - 1) / SP_WARPS), SP_WARPS * 32, smem, s>>>(a);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError());
    return 0;
}
template <int D>
int launch_space_mma(const SmallAttnArgs& a, cudaStream_t s) {
    return a.g == 1 ? launch_space_mma_impl<D, true>(a, s) : launch_space_mma_impl<D, false>(a, s);
}

static inline bool al16p(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

int d4_pool_attn_ok(const SmallAttnArgs& a) {
    // attention-residual pools: one query per token, 4 x 64 heads, unit-stride 256-wide rows
    return a.nq == 1 && a.g == 1 && a.hkv == 4 && a.d == 64 && !a.v0 && !a.belief && !a.mask_agent && a.softclamp <= 0.f && a.n >= 1 &&
           ((a.q_sb | a.k_sb | a.k_sj | a.v_sb | a.v_sj | a.out_sb) & 3) == 0 && al16p(a.q) && al16p(a.k) && al16p(a.v) && al16p(a.out) && al16p(a.k_gamma) &&
           (!a.gate_w || (al16p(a.gate_x) && al16p(a.gate_w) && (a.gate_x_ld & 3) == 0 && (a.gate_D & 3) == 0));
}

int d4_small_attn(const SmallAttnArgs& a, cudaStream_t s) {
    if (a.nb <= 0) return 0;
    if (a.gate_w && !d4_pool_attn_ok(a)) return d4_fail("small_attn: in-kernel gate logits are only implemented by the pool kernel");
    if (d4_pool_attn_ok(a)) {
        pool_attn_kernel<<<(unsigned)((a.nb + PL_WARPS - 1) / PL_WARPS), PL_WARPS * 32, 0, s>>>(a);
        D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError());
        return 0;
    }
    // space attention of one frame: S x S, S <= 16
    if (a.nq == a.n && a.n <= 16 && a.n >= 1 &&
        ((a.q_sb | a.q_si | a.k_sb | a.k_sj | a.v_sb | a.v_sj | a.v0_sb | a.v0_sj) & 3) == 0 && al16p(a.q) && al16p(a.k) && al16p(a.v) &&
        (!a.v0 || al16p(a.v0)) && al16p(a.k_gamma)) {
        static int space_v = -1;
        if (space_v < 0) { const char* v = getenv("D4_SPACE_V"); space_v = v ? atoi(v) : 3; }      // 2: the shared-memory staged version (cross-check)
        if (a.allow_tensor && space_v == 3 && d4_space_attn_reg_ok(a)) return d4_space_attn_reg(a, s);
        if (a.allow_tensor && a.d == 64 && (a.out_si & 1) == 0 && (a.out_sb & 1) == 0 && (reinterpret_cast<uintptr_t>(a.out) & 7) == 0)
            return launch_space_mma<64>(a, s);
        if (a.allow_tensor && a.d == 32 && (a.out_si & 1) == 0 && (a.out_sb & 1) == 0 && (reinterpret_cast<uintptr_t>(a.out) & 7) == 0)
            return launch_space_mma<32>(a, s);
        if (a.d == 64) return launch_space<64>(a, s);
        if (a.d == 32) return launch_space<32>(a, s);
        if (a.d == 16) return launch_space<16>(a, s);
        if (a.d == 128) return launch_space<128>(a, s);
    }
    if (a.n > 64 || a.n < 1) return d4_fail("small_attn: %d keys unsupported (1..64)", a.n);
    if (a.d % 4 != 0 || a.d > 128) return d4_fail("small_attn: head dim %d unsupported", a.d);
    if (a.belief && a.nq != a.n) return d4_fail("small_attn: belief projection needs nq == n");
    const size_t smem = (size_t)SA_WARPS * 2 * a.n * (a.d + 4) * sizeof(float);
    static size_t configured = 0;
    if (smem > configured) {
        D4_CUDA_OK(cudaFuncSetAttribute(small_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const long long items = (long long)a.nb * a.hkv;
    small_attn_kernel<<<(unsigned)((items + SA_WARPS - 1) / SA_WARPS), SA_WARPS * 32, smem, s>>>(a);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError());
    return 0;
}

int d4_time_attn(const TimeAttnArgs& a, cudaStream_t s) {
    if (a.M <= 0) return 0;
    if (a.t < 0 || a.t >= a.Tmax) return d4_fail("time_attn: t=%d outside the KV buffer (Tmax=%d)", a.t, a.Tmax);
    if (a.t > 32 * TA_MAXTILES) return d4_fail("time_attn: context %d > %d unsupported", a.t, 32 * TA_MAXTILES);
    if (a.variant == 1) return d4_time_attn_bulk(a, s);
#define D4_TA_CASE(DD, GG) if (a.d == DD && a.g == GG) return launch_time_attn<DD, GG>(a, s);
    D4_TA_CASE(64, 1) D4_TA_CASE(64, 2) D4_TA_CASE(32, 1) D4_TA_CASE(32, 2) D4_TA_CASE(16, 1) D4_TA_CASE(16, 2) D4_TA_CASE(128, 1)
#undef D4_TA_CASE
    return d4_fail("time_attn: (dim_head=%d, query groups=%d) has no kernel instantiation", a.d, a.g);
}
