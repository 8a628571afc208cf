// Tensor-core attention within one frame (the video tokenizer's encoder / decoder transformers, reference dreamer4.py:4360, 3655;
// arithmetic of reference dreamer4.py:1968-2075 + naive_attend 1683-1756, as frame_attn.cu states it).
//
// tf32x3 / f16x3 / tf32 engine modes, head dim 64, up to 128 keys: the 128 x 128 x 64 attention of a tokenizer
// frame is 4.2 MFLOP per (frame, head) - a third of the tokenizer's kernel time on exact-fp32 FMA (profiles/r2_tokenizer_launches.txt,
// 550 us per launch at 128 frames).  Same structure as above - one CTA per (frame, kv head), K and V staged once in shared memory, the
// 8 warps take 16-query tiles round-robin - with Q K^T and P V as 3-term TF32 mma.sync.m16n8k8 products (fp32-accurate), the scores of
// a tile (16 x n) and their softmax in registers, and the score C-fragments reused as the probability A-fragments (k-step s of P V
// covers keys 8s + {2t, 2t+1}), as in space_attn.cu.
//   fragment layouts: g = lane / 4, t = lane % 4
//   Q K^T contraction index: k-step s, fragment index t <-> head dim 16 t + 2 s, t + 4 <-> 16 t + 2 s + 1: a lane's Q values are the 16
//   contiguous floats [16 t, 16 t + 16) of its two rows (global -> registers, coalesced).  K is staged with its columns permuted so that
//   the matching pair of a lane is one conflict-free 8-byte shared load: dim 16 t + 2 s + e lives at column
//   4 (s & 3) + 32 (s >> 2) + 2 (t & 1) + 16 (t >> 1) + e  (pitch 68: the 16 lanes of a half-warp hit 32 distinct banks).
//   V keeps its natural columns: B-fragment (key 8 s + 2 t (+1), dim 8 n + g) is conflict-free at pitch 68 as well.
#include "kernels.h"
#include <float.h>
#include <stdlib.h>

namespace {

__device__ __forceinline__ float lerp_(float a, float b, float w) {      // torch.lerp as ATen evaluates it
    const float d = b - a;
    return (w < 0.5f) ? a + w * d : b - d * (1.f - w);
}

constexpr int FM_WARPS = 8, FM_D = 64, FM_P = 68, FM_MAXN = 128;

__device__ __forceinline__ void fm_mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void fm_split(float x, uint32_t& hi, uint32_t& lo) { tf32_split_mma(x, hi, lo); }
__device__ __forceinline__ void fm_mma_3x(float (&c)[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4], const uint32_t (&bhi)[2], const uint32_t (&blo)[2]) {
    fm_mma_tf32(c, alo, bhi);
    fm_mma_tf32(c, ahi, blo);
    fm_mma_tf32(c, ahi, bhi);
}
__device__ __forceinline__ int fm_kcol(int dim) {          // staged column of head dim `dim` of a key row
    const int t = dim >> 4, j = dim & 15, s = j >> 1, e = j & 1;
    return 4 * (s & 3) + 32 * (s >> 2) + 2 * (t & 1) + 16 * (t >> 1) + e;
}

__device__ __forceinline__ int fm_kdim(int col) {          // inverse of fm_kcol
    const int t = (((col >> 4) & 1) << 1) | ((col >> 1) & 1), s = (((col >> 5) & 1) << 2) | ((col >> 2) & 3);
    return 16 * t + 2 * s + (col & 1);
}

template <int NT, int MINB>        // key tiles of 8: n <= 8 * NT; MINB resident CTAs per SM asked of the register allocator
__global__ void __launch_bounds__(FM_WARPS * 32, MINB) frame_attn_mma_kernel(SmallAttnArgs a) {
    extern __shared__ __align__(16) float smem[];
    constexpr int D = FM_D, P = FM_P, NP = 8 * NT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x / a.hkv, hk = blockIdx.x % a.hkv;
    const int n = a.n, ns = a.mask_agent;
    const int g = lane >> 2, t = lane & 3;
    float* Ks = smem;                      // [NP][P], columns permuted (fm_kcol), normalised keys
    float* Vs = Ks + NP * P;               // [NP][P], value-residual lerp applied
    float* vinv = Vs + NP * P;             // [NP] 1 / |v_j| (belief projection)

    // stage K (columns permuted) and V (value-residual lerp applied) of this (frame, head): 4 float4 per thread in flight
    constexpr int STG = 4;
    for (int base = 0; base < NP * (D / 4); base += STG * FM_WARPS * 32) {
        float4 kv[STG], vv[STG], rv[STG]; float mw[STG];
#pragma unroll
        for (int u = 0; u < STG; ++u) {
            const int idx = base + u * FM_WARPS * 32 + threadIdx.x, j = idx / (D / 4), c = (idx % (D / 4)) * 4;
            kv[u] = vv[u] = rv[u] = make_float4(0.f, 0.f, 0.f, 0.f); mw[u] = 0.f;
            if (idx < NP * (D / 4) && j < n) {
                kv[u] = __ldg(reinterpret_cast<const float4*>(a.k + b * a.k_sb + (long long)j * a.k_sj + (long long)hk * D + c));
                vv[u] = __ldg(reinterpret_cast<const float4*>(a.v + b * a.v_sb + (long long)j * a.v_sj + (long long)hk * D + c));
                if (a.v0) {
                    rv[u] = __ldg(reinterpret_cast<const float4*>(a.v0 + b * a.v0_sb + (long long)j * a.v0_sj + (long long)hk * D + c));
                    mw[u] = __ldg(a.mix + b * a.mix_sb + (long long)j * a.mix_sj + hk);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < STG; ++u) {
            const int idx = base + u * FM_WARPS * 32 + threadIdx.x, j = idx / (D / 4), c = (idx % (D / 4)) * 4;
            if (idx >= NP * (D / 4)) continue;
            if (a.v0 && j < n) {
                const float w = sigmoidf_(mw[u]);
                vv[u].x = lerp_(vv[u].x, rv[u].x, w); vv[u].y = lerp_(vv[u].y, rv[u].y, w); vv[u].z = lerp_(vv[u].z, rv[u].z, w); vv[u].w = lerp_(vv[u].w, rv[u].w, w);
            }
            *reinterpret_cast<float2*>(Ks + j * P + fm_kcol(c)) = make_float2(kv[u].x, kv[u].y);
            *reinterpret_cast<float2*>(Ks + j * P + fm_kcol(c + 2)) = make_float2(kv[u].z, kv[u].w);
            *reinterpret_cast<float4*>(Vs + j * P + c) = vv[u];
        }
    }
    __syncthreads();
    // MultiHeadRMSNorm on keys (reference dreamer4.py:1663-1679) and the value norms of the belief projection: a warp per row, a lane
    // per two columns (conflict-free), sums in lane order then a shuffle tree
    const float sqrt_d = 8.f;
    {
        const int c0 = lane, c1 = lane + 32;
        const float g0 = (a.k_gamma[hk * D + fm_kdim(c0)] + 1.f) * sqrt_d, g1 = (a.k_gamma[hk * D + fm_kdim(c1)] + 1.f) * sqrt_d;
        for (int j = warp; j < NP; j += FM_WARPS) {
            float* kr = Ks + j * P;
            const float* vr = Vs + j * P;
            const float k0 = kr[c0], k1 = kr[c1], v0 = vr[c0], v1 = vr[c1];
            const float ss = warp_sum(fmaf(k1, k1, k0 * k0)), vs = warp_sum(fmaf(v1, v1, v0 * v0));
            const float inv = 1.f / fmaxf(sqrtf(ss), D4_L2_EPS);
            kr[c0] = (k0 * inv) * g0; kr[c1] = (k1 * inv) * g1;
            if (lane == 0) vinv[j] = 1.f / fmaxf(sqrtf(vs), D4_L2_EPS);
        }
    }
    __syncthreads();

    for (int gi = 0; gi < a.g; ++gi) {
        const int hq = hk * a.g + gi;
        for (int i0 = 16 * warp; i0 < a.nq; i0 += 16 * FM_WARPS) {
            // ---- Q fragments of rows i0 + g, i0 + g + 8: dims [16 t, 16 t + 16)
            float qf[2][16];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = i0 + g + 8 * h;
                const float* qp = a.q + b * a.q_sb + (long long)i * a.q_si + (long long)hq * D + 16 * t;
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const float4 v = (i < a.nq) ? __ldg(reinterpret_cast<const float4*>(qp + 4 * q4)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    qf[h][4 * q4] = v.x; qf[h][4 * q4 + 1] = v.y; qf[h][4 * q4 + 2] = v.z; qf[h][4 * q4 + 3] = v.w;
                }
            }
            // ---- scores: 16 queries x NP keys
            float sc[NT][4];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int r = 0; r < 4; ++r) sc[nt][r] = 0.f;
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                uint32_t ahi[4], alo[4];
                fm_split(qf[0][2 * s], ahi[0], alo[0]);
                fm_split(qf[1][2 * s], ahi[1], alo[1]);
                fm_split(qf[0][2 * s + 1], ahi[2], alo[2]);
                fm_split(qf[1][2 * s + 1], ahi[3], alo[3]);
                const int kc = 4 * (s & 3) + 32 * (s >> 2) + 2 * (t & 1) + 16 * (t >> 1);
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const float2 kk = *reinterpret_cast<const float2*>(Ks + (8 * nt + g) * P + kc);
                    uint32_t bhi[2], blo[2];
                    fm_split(kk.x, bhi[0], blo[0]);
                    fm_split(kk.y, bhi[1], blo[1]);
                    fm_mma_3x(sc[nt], ahi, alo, bhi, blo);
                }
            }
            // ---- scale, softclamp, special-token mask, softmax over the keys of each row (row g: regs 0, 1; row g + 8: regs 2, 3)
            if (a.softclamp > 0.f) {          // tanh(s / c) c  (reference dreamer4.py:1723-1724)
                const float pre = a.scale * (1.f / a.softclamp);
                float amax = 0.f;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int r = 0; r < 4; ++r) { sc[nt][r] *= pre; amax = fmaxf(amax, fabsf(sc[nt][r])); }
                if (__all_sync(D4_FULL, amax <= D4_TANH_POLY_MAX)) {
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int r = 0; r < 4; ++r) sc[nt][r] = tanh_small_(sc[nt][r]) * a.softclamp;
                } else {
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int r = 0; r < 4; ++r) sc[nt][r] = tanhf(sc[nt][r]) * a.softclamp;
                }
            } else {
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int r = 0; r < 4; ++r) sc[nt][r] *= a.scale;
            }
            if (ns > 0 && i0 < a.nq - ns) {          // queries i < nq - ns do not see the last ns keys (1769-1783)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int i = i0 + g + ((r & 2) ? 8 : 0), j = 8 * nt + 2 * t + (r & 1);
                        if (i < a.nq - ns && j >= n - ns) sc[nt][r] = -FLT_MAX;
                    }
            }
            if (n < NP) {                            // key padding of the last tiles
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int r = 0; r < 4; ++r) if (8 * nt + 2 * t + (r & 1) >= n) sc[nt][r] = -INFINITY;
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float mx = -INFINITY;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) mx = fmaxf(mx, fmaxf(sc[nt][2 * half], sc[nt][2 * half + 1]));
                mx = fmaxf(mx, __shfl_xor_sync(D4_FULL, mx, 1)); mx = fmaxf(mx, __shfl_xor_sync(D4_FULL, mx, 2));
                float sum = 0.f;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const float e0 = __expf(sc[nt][2 * half] - mx), e1 = __expf(sc[nt][2 * half + 1] - mx);
                    sc[nt][2 * half] = e0; sc[nt][2 * half + 1] = e1; sum += e0 + e1;
                }
                sum += __shfl_xor_sync(D4_FULL, sum, 1); sum += __shfl_xor_sync(D4_FULL, sum, 2);
                const float inv = 1.f / sum;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) { sc[nt][2 * half] *= inv; sc[nt][2 * half + 1] *= inv; }
            }
            // ---- out = P V: k-step s = key tile s (fragment index t <-> key 8 s + 2 t, t + 4 <-> key 8 s + 2 t + 1), 8 column tiles
            float o[8][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int r = 0; r < 4; ++r) o[nt][r] = 0.f;
#pragma unroll
            for (int s = 0; s < NT; ++s) {
                uint32_t phi[4], plo[4];
                fm_split(sc[s][0], phi[0], plo[0]);
                fm_split(sc[s][2], phi[1], plo[1]);
                fm_split(sc[s][1], phi[2], plo[2]);
                fm_split(sc[s][3], phi[3], plo[3]);
                const float* v0r = Vs + (8 * s + 2 * t) * P + g;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    uint32_t bhi[2], blo[2];
                    fm_split(v0r[8 * nt], bhi[0], blo[0]);
                    fm_split(v0r[P + 8 * nt], bhi[1], blo[1]);
                    fm_mma_3x(o[nt], phi, plo, bhi, blo);
                }
            }
            // lane (g, t) holds out[i0 + g + 8 h][8 nt + 2 t + {0, 1}] = o[nt][2 h + {0, 1}]
            // ---- belief projection (self-attention: key i is token i), head gate, store
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = i0 + g + 8 * h;
                const bool ok = i < a.nq;
                if (a.belief) {
                    const int iv = ok ? i : 0;
                    const float vi = vinv[iv];
                    float dot = 0.f;
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt) {
                        const float2 vv = *reinterpret_cast<const float2*>(Vs + iv * P + 8 * nt + 2 * t);
                        dot = fmaf(o[nt][2 * h], vv.x * vi, dot); dot = fmaf(o[nt][2 * h + 1], vv.y * vi, dot);
                    }
                    dot += __shfl_xor_sync(D4_FULL, dot, 1); dot += __shfl_xor_sync(D4_FULL, dot, 2);
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt) {
                        const float2 vv = *reinterpret_cast<const float2*>(Vs + iv * P + 8 * nt + 2 * t);
                        o[nt][2 * h] -= dot * (vv.x * vi); o[nt][2 * h + 1] -= dot * (vv.y * vi);
                    }
                }
                if (ok) {
                    const float gate = a.gate ? sigmoidf_(a.gate[b * a.gate_sb + (long long)i * a.gate_si + hq]) : 1.f;
                    float* op = a.out + b * a.out_sb + (long long)i * a.out_si + (long long)hq * D + 2 * t;
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt) *reinterpret_cast<float2*>(op + 8 * nt) = make_float2(o[nt][2 * h] * gate, o[nt][2 * h + 1] * gate);
                }
            }
        }
    }
}

template <int NT, int MINB>
int launch_fm(const SmallAttnArgs& a, cudaStream_t s) {
    const size_t smem = ((size_t)2 * 8 * NT * FM_P + 8 * NT) * sizeof(float);
    static bool configured = false;
    if (!configured) {
        D4_CUDA_OK(cudaFuncSetAttribute(frame_attn_mma_kernel<NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    frame_attn_mma_kernel<NT, MINB><<<(unsigned)((long long)a.nb * a.hkv), FM_WARPS * 32, smem, s>>>(a);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace

int d4_frame_attn_mma_ok(const SmallAttnArgs& a) {
    static const bool enabled = [] { const char* e = getenv("D4_FRAME_MMA"); return !(e && e[0] == '0'); }();      // D4_FRAME_MMA=0: FMA kernel (A/B runs)
    // (grouped queries - a.g > 1 - are written out in the kernel but no test reaches them: they stay on the FMA kernel, which the simulator covers)
    return enabled && a.allow_tensor && a.g == 1 && a.d == FM_D && a.n >= 1 && a.n <= FM_MAXN && a.nq >= 1 && (!a.belief || a.nq == a.n) &&
           ((a.q_sb | a.q_si | a.k_sb | a.k_sj | a.v_sb | a.v_sj | a.v0_sb | a.v0_sj) & 3) == 0 && ((a.out_sb | a.out_si) & 1) == 0 &&
           (reinterpret_cast<uintptr_t>(a.q) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.k) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.v) & 15) == 0 &&
           (!a.v0 || ((reinterpret_cast<uintptr_t>(a.v0) & 15) == 0 && a.mix)) && (reinterpret_cast<uintptr_t>(a.out) & 7) == 0 ? 1 : 0;
}

int d4_frame_attn_mma(const SmallAttnArgs& a, cudaStream_t s) {
    if (!d4_frame_attn_mma_ok(a)) return d4_fail("frame_attn_mma: shape / alignment not supported");
    static const int minb = [] { const char* e = getenv("D4_FRAME_MINB"); return e ? atoi(e) : 2; }();       // 2 CTAs per SM: 128 registers (16 bytes spilled), 25.8k vs 24.9k tokenizer frames/s
    if (a.n <= 32) return launch_fm<4, 2>(a, s);
    if (a.n <= 64) return launch_fm<8, 1>(a, s);
    if (a.n <= 96) return launch_fm<12, 1>(a, s);
    return minb == 2 ? launch_fm<16, 2>(a, s) : launch_fm<16, 1>(a, s);
}
