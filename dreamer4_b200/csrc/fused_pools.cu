// Fused learned-query pools on either side of the transformer (reference dreamer4.py:2179-2210, built at 4822-4834):
//
//   l2s_fused_kernel  latents (B, N, Dl) -> gated attention output (B, nsp, Dq) of `latents_to_spatial_tokens`
//                     (the to_out projection stays a GEMM).  Replaces  row_rstd + GEMM(Dl -> 2*Dkv) + attention:  the
//                     (B*N, 2*Dkv) key/value tensor (537 MB at config 4) is never materialised.  Keys are projected in
//                     registers (they are needed in full for the per-head key RMSNorm); values are NOT projected per key:
//                     sum_j p_ij W_v x_j == W_v (sum_j p_ij x_j), so the probabilities pool the Dl-wide normalised
//                     latents and W_v is applied once per (query, head).
//   lp_fused_kernel   spatial-token keys/values (B, nsp, 2*Dkv) -> predicted latents (B, N, Dl) of `to_latent_pred`.
//                     Replaces  attention (N queries x nsp keys) + GEMM(Dq -> Dl) and the (B*N, Dq) tensor between them:
//                     with only nsp keys,  pred_i = sum_{j,h} gate_ih p_ijh (W_comb,h v_jh):  the nsp*hq vectors W_comb,h v_jh
//                     are formed once per frame and every query just mixes them.
//
// Both are exact fp32 re-associations of the reference arithmetic (no reduced precision anywhere).
#include "kernels.h"
#include <float.h>
#include <stdlib.h>

namespace {

constexpr int L2S_MAXQ = 8;        // learned queries x query groups per kv head
constexpr int L2S_MAXN = 64;       // latent tokens

// MMA = true (tensor-core engine modes): the key projection K_h = X W_k,h^T (N x d, reduction Dl) runs as m16n8k8 TF32
// mma.sync tiles with the 3-term split (fp32-accurate), the key norms and q.k dots are reduced straight from the
// accumulator fragments; MMA = false keeps it in exact-fp32 FMA.  Fragment layouts: see space_attn_mma_kernel (attn.cu).
__device__ __forceinline__ void l2s_mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void l2s_split(float x, uint32_t& hi, uint32_t& lo) {      // (tf32_split_mma of common.cuh was measured here: 255 vs 226 us per launch - cvt.rna kept)
    const float h = tf32_rna(x);
    hi = __float_as_uint(h);
    lo = __float_as_uint(tf32_rna(x - h));
}

template <int DL, bool MMA, int NQC>
__global__ void __launch_bounds__(256) l2s_fused_kernel(L2sArgs a) {
    // one CTA per frame, one warp per kv head
    constexpr int XP = DL + 4;                         // row pitch of the latent / weight tiles (conflict-free fragment loads)
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int b = blockIdx.x;
    const int N = a.N, d = a.d, NQ = a.nsp * a.g;
    float* xs = smem;                                  // [64][XP] normalised latents (rows N.. are zero)
    float* pw_all = xs + L2S_MAXN * XP;                // per warp [L2S_MAXN][NQC]
    float* zs_all = pw_all + nwarps * L2S_MAXN * NQC;   // per warp [NQC][DL]
    float* qs_all = zs_all + nwarps * NQC * DL;   // per warp [NQC][d]  queries x key gain
    float* ws_all = qs_all + nwarps * NQC * d;    // per warp [d][XP]        this head's key (then value) projection rows

    // ---- normalised latents of this frame: x * rsqrt(mean(x^2) + eps)  (norm_context gamma is folded into w_k / w_v)
    const float* xb = a.latent + (long long)b * N * DL;
    for (int j = warp; j < L2S_MAXN; j += nwarps) {
        float v[(DL + 31) / 32], ss = 0.f;
#pragma unroll
        for (int e = 0; e < (DL + 31) / 32; ++e) { const int c = lane + 32 * e; v[e] = (c < DL && j < N) ? xb[j * DL + c] : 0.f; ss += v[e] * v[e]; }
        ss = warp_sum(ss);
        const float r = rsqrtf(ss / (float)DL + D4_RMS_EPS);
#pragma unroll
        for (int e = 0; e < (DL + 31) / 32; ++e) { const int c = lane + 32 * e; if (c < DL) xs[j * XP + c] = v[e] * r; }
    }
    __syncthreads();

    for (int hk = warp; hk < a.h; hk += nwarps) {
        float* pw = pw_all + warp * L2S_MAXN * NQC;
        float* zs = zs_all + warp * NQC * DL;
        float* qs = qs_all + warp * NQC * d;
        float* wsm = ws_all + warp * d * XP;
        const float sqrt_d = sqrtf((float)d);
        {   // stage W_k,h: one coalesced pass instead of d * DL / 4 uniform global loads inside the score loop
            const float4* src = reinterpret_cast<const float4*>(a.w_k + (long long)hk * d * DL);
            for (int idx = lane; idx < d * DL / 4; idx += 32) {
                const int c = idx / (DL / 4), e = (idx % (DL / 4)) * 4;
                *reinterpret_cast<float4*>(wsm + c * XP + e) = __ldg(src + idx);
            }
        }
        for (int idx = lane; idx < NQ * d; idx += 32) {
            const int qi = idx / d, c = idx - qi * d;
            const int i = qi / a.g, gi = qi - i * a.g;
            qs[idx] = a.q[(long long)i * a.Dq + (hk * a.g + gi) * d + c] * ((a.k_gamma[hk * d + c] + 1.f) * sqrt_d);
        }
        __syncwarp();
        float sc[2][NQC];
        if (MMA) {
            // ---- K_h = X W_k,h^T on mma.sync: per 16-key tile an accumulator of d / 8 column tiles; |k|^2 and the q.k dots are
            //      reduced from the fragments (rows g and g + 8 of the tile, columns 2t, 2t + 1), raw scores parked in pw
            const int g = lane >> 2, t = lane & 3;
            for (int mt = 0; mt * 16 < N; ++mt) {
                float ss[2] = {0.f, 0.f}, dot[2][NQC];
#pragma unroll
                for (int qi = 0; qi < NQC; ++qi) { dot[0][qi] = 0.f; dot[1][qi] = 0.f; }
                uint32_t ahi[DL / 8][4], alo[DL / 8][4];
#pragma unroll
                for (int ks = 0; ks < DL / 8; ++ks) {
                    const float* xr = xs + (mt * 16 + g) * XP + ks * 8 + t;
                    l2s_split(xr[0], ahi[ks][0], alo[ks][0]); l2s_split(xr[8 * XP], ahi[ks][1], alo[ks][1]);
                    l2s_split(xr[4], ahi[ks][2], alo[ks][2]); l2s_split(xr[8 * XP + 4], ahi[ks][3], alo[ks][3]);
                }
                for (int nt = 0; nt * 8 < d; ++nt) {
                    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int ks = 0; ks < DL / 8; ++ks) {
                        uint32_t bhi[2], blo[2];
                        const float* wr = wsm + (nt * 8 + g) * XP + ks * 8 + t;
                        l2s_split(wr[0], bhi[0], blo[0]); l2s_split(wr[4], bhi[1], blo[1]);
                        l2s_mma_tf32(acc, alo[ks], bhi); l2s_mma_tf32(acc, ahi[ks], blo); l2s_mma_tf32(acc, ahi[ks], bhi);
                    }
                    ss[0] = fmaf(acc[0], acc[0], fmaf(acc[1], acc[1], ss[0]));
                    ss[1] = fmaf(acc[2], acc[2], fmaf(acc[3], acc[3], ss[1]));
#pragma unroll
                    for (int qi = 0; qi < NQC; ++qi) {
                        if (qi < NQ) {
                            const float2 qv = *reinterpret_cast<const float2*>(qs + qi * d + nt * 8 + 2 * t);
                            dot[0][qi] = fmaf(acc[0], qv.x, fmaf(acc[1], qv.y, dot[0][qi]));
                            dot[1][qi] = fmaf(acc[2], qv.x, fmaf(acc[3], qv.y, dot[1][qi]));
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    ss[r] += __shfl_xor_sync(D4_FULL, ss[r], 1); ss[r] += __shfl_xor_sync(D4_FULL, ss[r], 2);
                    const float inv = a.scale / fmaxf(sqrtf(ss[r]), D4_L2_EPS);
                    const int key = mt * 16 + g + 8 * r;
#pragma unroll
                    for (int qi = 0; qi < NQC; ++qi) {
                        if (qi < NQ) {
                            float dv = dot[r][qi];
                            dv += __shfl_xor_sync(D4_FULL, dv, 1); dv += __shfl_xor_sync(D4_FULL, dv, 2);
                            if (t == 0) pw[key * NQC + qi] = dv * inv;
                        }
                    }
                }
            }
            __syncwarp();
#pragma unroll
            for (int kk = 0; kk < 2; ++kk)
#pragma unroll
                for (int qi = 0; qi < NQC; ++qi) {
                    const int j = lane + 32 * kk;
                    sc[kk][qi] = (j < N && qi < NQ) ? pw[j * NQC + qi] : -INFINITY;
                }
            __syncwarp();
        } else {
        // ---- scores: lane = key (two key slots), keys projected on the fly, |k| accumulated alongside
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
            const int j = lane + 32 * kk;
            float xr[DL];
#pragma unroll
            for (int e = 0; e < DL; e += 4) {
                const float4 t = (j < N) ? *reinterpret_cast<const float4*>(xs + j * XP + e) : make_float4(0.f, 0.f, 0.f, 0.f);
                xr[e] = t.x; xr[e + 1] = t.y; xr[e + 2] = t.z; xr[e + 3] = t.w;
            }
            float ss = 0.f, dot[NQC];
#pragma unroll
            for (int qi = 0; qi < NQC; ++qi) dot[qi] = 0.f;
            if (kk * 32 < N) {                                  // warp-uniform
                for (int c = 0; c < d; ++c) {
                    float k = 0.f;
#pragma unroll
                    for (int e = 0; e < DL; e += 4) {
                        const float4 w = *reinterpret_cast<const float4*>(wsm + c * XP + e);
                        k = fmaf(xr[e], w.x, k); k = fmaf(xr[e + 1], w.y, k); k = fmaf(xr[e + 2], w.z, k); k = fmaf(xr[e + 3], w.w, k);
                    }
                    ss = fmaf(k, k, ss);
#pragma unroll
                    for (int qi = 0; qi < NQC; ++qi) if (qi < NQ) dot[qi] = fmaf(qs[qi * d + c], k, dot[qi]);
                }
            }
            const float inv = a.scale / fmaxf(sqrtf(ss), D4_L2_EPS);
#pragma unroll
            for (int qi = 0; qi < NQC; ++qi) sc[kk][qi] = (j < N) ? dot[qi] * inv : -INFINITY;
        }
        }
        // ---- softmax over the N keys of every query
#pragma unroll
        for (int qi = 0; qi < NQC; ++qi) {
            if (qi < NQ) {
                const float mx = warp_max(fmaxf(sc[0][qi], sc[1][qi]));
                const float e0 = (lane < N) ? expf(sc[0][qi] - mx) : 0.f, e1 = (lane + 32 < N) ? expf(sc[1][qi] - mx) : 0.f;
                const float inv = 1.f / warp_sum(e0 + e1);
                pw[lane * NQC + qi] = e0 * inv;
                pw[(lane + 32) * NQC + qi] = e1 * inv;
            }
        }
        __syncwarp();
        // ---- pooled normalised latents z_q = sum_j p_qj x_j : lane = latent channel
#pragma unroll
        for (int eb = 0; eb < DL; eb += 32) {
            const int e = eb + lane;
            float z[NQC];
#pragma unroll
            for (int qi = 0; qi < NQC; ++qi) z[qi] = 0.f;
            if (e < DL) {
                for (int j = 0; j < N; ++j) {
                    const float xv = xs[j * XP + e];
#pragma unroll
                    for (int q4 = 0; q4 < NQC; q4 += 4) {
                        const float4 p4 = *reinterpret_cast<const float4*>(pw + j * NQC + q4);
                        z[q4] = fmaf(p4.x, xv, z[q4]); z[q4 + 1] = fmaf(p4.y, xv, z[q4 + 1]);
                        z[q4 + 2] = fmaf(p4.z, xv, z[q4 + 2]); z[q4 + 3] = fmaf(p4.w, xv, z[q4 + 3]);
                    }
                }
#pragma unroll
                for (int qi = 0; qi < NQC; ++qi) zs[qi * DL + e] = z[qi];
            }
        }
        __syncwarp();
        // ---- values: o_q = W_v,h z_q, head gate, store : lane = head channel
        {
            const float4* src = reinterpret_cast<const float4*>(a.w_v + (long long)hk * d * DL);
            for (int idx = lane; idx < d * DL / 4; idx += 32) {
                const int c = idx / (DL / 4), e = (idx % (DL / 4)) * 4;
                *reinterpret_cast<float4*>(wsm + c * XP + e) = __ldg(src + idx);
            }
        }
        __syncwarp();
        for (int c = lane; c < d; c += 32) {
            float o[NQC];
#pragma unroll
            for (int qi = 0; qi < NQC; ++qi) o[qi] = 0.f;
#pragma unroll
            for (int e = 0; e < DL; e += 4) {
                // row c of W_v: lanes read different rows -> rotate the chunk order so the 32 rows hit distinct banks
                const int er = e;          // rows are XP = DL + 4 floats apart: lanes reading different rows hit different banks
                const float4 w = *reinterpret_cast<const float4*>(wsm + c * XP + er);
#pragma unroll
                for (int qi = 0; qi < NQC; ++qi) {
                    if (qi < NQ) {
                        const float4 z = *reinterpret_cast<const float4*>(zs + qi * DL + er);
                        o[qi] = fmaf(w.x, z.x, o[qi]); o[qi] = fmaf(w.y, z.y, o[qi]); o[qi] = fmaf(w.z, z.z, o[qi]); o[qi] = fmaf(w.w, z.w, o[qi]);
                    }
                }
            }
#pragma unroll
            for (int qi = 0; qi < NQC; ++qi) {
                if (qi < NQ) {
                    const int i = qi / a.g, gi = qi - i * a.g, hq = hk * a.g + gi;
                    const float gate = sigmoidf_(a.gate[i * a.hq + hq]);
                    a.out[((long long)b * a.nsp + i) * a.Dq + hq * d + c] = o[qi] * gate;
                }
            }
        }
        __syncwarp();       // pw / zs / qs are reused by this warp's next head
    }
}

// -------------------------------------------------------------------------------------------------
constexpr int LP_MAXJH = 64;       // nsp * hq mixing vectors

__global__ void __launch_bounds__(256) lp_fused_kernel(LpArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, nthr = blockDim.x, warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
    const int b = blockIdx.x;
    const int nsp = a.nsp, h = a.h, hq = a.hq, g = hq / h, d = a.d, Dkv = h * d, Dq = hq * d, N = a.N, Dl = a.Dl;
    const int dp = d + 4;                                   // padded head pitch: conflict-free float4 reads across heads
    const int JH = nsp * hq;
    float* ks = smem;                                       // [nsp][h][dp]   keys x key gain / |k|
    float* vs = ks + nsp * h * dp;                          // [nsp][h][dp]
    float* ys = vs + nsp * h * dp;                          // [JH][Dl]       W_comb,hq v_(j,hk)
    float* cf = ys + LP_MAXJH * Dl;                         // [N][JH + 1]    gate * softmax probabilities

    const float* kvb = a.kv + (long long)b * nsp * 2 * Dkv;
    for (int idx = tid; idx < nsp * Dkv; idx += nthr) {
        const int j = idx / Dkv, r = idx - j * Dkv, hk = r / d, c = r - hk * d;
        ks[(j * h + hk) * dp + c] = kvb[(long long)j * 2 * Dkv + r];
        vs[(j * h + hk) * dp + c] = kvb[(long long)j * 2 * Dkv + Dkv + r];
    }
    __syncthreads();
    // key RMSNorm per (key, head): k <- l2norm(k) * (gamma + 1) * sqrt(d)
    const float sqrt_d = sqrtf((float)d);
    for (int jh = warp; jh < nsp * h; jh += nwarps) {
        float* kr = ks + jh * dp;
        const int hk = jh % h;
        float ss = 0.f;
        for (int c = lane; c < d; c += 32) ss += kr[c] * kr[c];
        const float inv = 1.f / fmaxf(sqrtf(warp_sum(ss)), D4_L2_EPS);
        for (int c = lane; c < d; c += 32) kr[c] = kr[c] * inv * ((a.k_gamma[hk * d + c] + 1.f) * sqrt_d);
    }
    // mixing vectors y[(j,hq)][e] = sum_c W_comb[e][hq*d + c] * v[j][hk][c]
    for (int o = tid; o < JH * Dl; o += nthr) {
        const int jh = o / Dl, e = o - jh * Dl, j = jh / hq, q = jh - j * hq, hk = q / g;
        const float* w = a.w_comb + (long long)e * Dq + q * d;
        const float* v = vs + (j * h + hk) * dp;
        float acc = 0.f;
        for (int c = 0; c < d; c += 4) {
            const float4 wv = __ldg(reinterpret_cast<const float4*>(w + c));
            const float4 vv = *reinterpret_cast<const float4*>(v + c);
            acc = fmaf(wv.x, vv.x, acc); acc = fmaf(wv.y, vv.y, acc); acc = fmaf(wv.z, vv.z, acc); acc = fmaf(wv.w, vv.w, acc);
        }
        ys[jh * Dl + e] = acc;
    }
    __syncthreads();
    // scores + softmax over the nsp keys for every (query, query head); coefficient = gate * probability
    for (int p = tid; p < N * hq; p += nthr) {
        const int i = p / hq, q = p - i * hq, hk = q / g;
        const float* qr = a.q + (long long)i * Dq + q * d;
        float s[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] = 0.f;
        for (int c = 0; c < d; c += 4) {
            const float4 qv = __ldg(reinterpret_cast<const float4*>(qr + c));
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (j < nsp) {
                    const float4 kv = *reinterpret_cast<const float4*>(ks + (j * h + hk) * dp + c);
                    s[j] = fmaf(qv.x, kv.x, s[j]); s[j] = fmaf(qv.y, kv.y, s[j]); s[j] = fmaf(qv.z, kv.z, s[j]); s[j] = fmaf(qv.w, kv.w, s[j]);
                }
            }
        }
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j) if (j < nsp) { s[j] *= a.scale; mx = fmaxf(mx, s[j]); }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) if (j < nsp) { s[j] = expf(s[j] - mx); sum += s[j]; }
        const float gate = sigmoidf_(a.gate[i * hq + q]) / sum;
#pragma unroll
        for (int j = 0; j < 8; ++j) if (j < nsp) cf[i * (JH + 1) + j * hq + q] = s[j] * gate;
    }
    __syncthreads();
    // pred[i][e] = sum_jh cf[i][jh] * y[jh][e]
    float* outb = a.pred + (long long)b * N * Dl;
    for (int o = tid; o < N * Dl; o += nthr) {
        const int i = o / Dl, e = o - i * Dl;
        float acc = 0.f;
        for (int jh = 0; jh < JH; ++jh) acc = fmaf(cf[i * (JH + 1) + jh], ys[jh * Dl + e], acc);
        outb[o] = acc;
    }
}


// Persistent version for large batches (head dim 64, N * hq <= 512 query rows): the learned queries (N x hq*d = 128 KB) and W_comb
// (Dl x hq*d = 64 KB) do not depend on the frame, yet lp_fused_kernel re-reads both from L2 for each of its B CTAs (393 MB of
// L2 -> SM traffic at 2048 frames, latency-bound loads inside the dot products: 289 us per launch for 1 GFLOP).  Here one CTA per SM
// keeps W_comb in shared memory and each thread keeps ITS query row (64 floats) in registers, then walks over its frames.  Every
// output is accumulated in the same order as in lp_fused_kernel, so the two are bit-identical (tests/test_gpu_parity.py).
constexpr int LPP_THREADS = 512;

template <int D>
__global__ void __launch_bounds__(LPP_THREADS, 1) lp_fused_persist_kernel(LpArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, nthr = LPP_THREADS, warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
    const int nsp = a.nsp, h = a.h, hq = a.hq, g = hq / h, Dkv = h * D, Dq = hq * D, N = a.N, Dl = a.Dl;
    constexpr int dp = D + 4;
    const int JH = nsp * hq, WP = Dq + 4;                   // W_comb row pitch: rows e, e + 1, .. of a warp start 4 banks apart
    float* wc = smem;                                       // [Dl][WP]
    float* ks = wc + (size_t)Dl * WP;                       // [nsp][h][dp]
    float* vs = ks + nsp * h * dp;
    float* ys = vs + nsp * h * dp;                          // [JH][Dl]
    float* cf = ys + LP_MAXJH * Dl;                         // [N][JH + 1]

    for (int idx = tid; idx < Dl * (Dq / 4); idx += nthr) {
        const int e = idx / (Dq / 4), c = (idx % (Dq / 4)) * 4;
        *reinterpret_cast<float4*>(wc + (size_t)e * WP + c) = __ldg(reinterpret_cast<const float4*>(a.w_comb + (long long)e * Dq + c));
    }
    // this thread's (query, query head): its row of the projected queries and its gate, for every frame
    // (query head)-major: the 32 lanes of a warp share a head, so their key reads are ONE broadcast 16-byte access instead of four
    // quarter-warp wavefronts (ncu: the first mapping spent 55 % of the shared-memory pipe's wavefronts here)
    const int my_q = tid / N, my_i = tid - my_q * N, my_hk = my_q / g;
    const bool has_pair = tid < N * hq;
    float qreg[D];
    float my_gate = 0.f;
    if (has_pair) {
        const float* qr = a.q + (long long)my_i * Dq + my_q * D;
#pragma unroll
        for (int c = 0; c < D; c += 4) {
            const float4 qv = __ldg(reinterpret_cast<const float4*>(qr + c));
            qreg[c] = qv.x; qreg[c + 1] = qv.y; qreg[c + 2] = qv.z; qreg[c + 3] = qv.w;
        }
        my_gate = sigmoidf_(a.gate[my_i * hq + my_q]);
    }
    const float sqrt_d = sqrtf((float)D);

    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        __syncthreads();                                    // the previous frame's cf / ys are no longer read (first pass: wc is staged)
        const float* kvb = a.kv + (long long)b * nsp * 2 * Dkv;
        for (int idx = tid; idx < nsp * Dkv; idx += nthr) {
            const int j = idx / Dkv, r = idx - j * Dkv, hk = r / D, c = r - hk * D;
            ks[(j * h + hk) * dp + c] = kvb[(long long)j * 2 * Dkv + r];
            vs[(j * h + hk) * dp + c] = kvb[(long long)j * 2 * Dkv + Dkv + r];
        }
        __syncthreads();
        for (int jh = warp; jh < nsp * h; jh += nwarps) {
            float* kr = ks + jh * dp;
            const int hk = jh % h;
            float ss = 0.f;
            for (int c = lane; c < D; c += 32) ss += kr[c] * kr[c];
            const float inv = 1.f / fmaxf(sqrtf(warp_sum(ss)), D4_L2_EPS);
            for (int c = lane; c < D; c += 32) kr[c] = kr[c] * inv * ((a.k_gamma[hk * D + c] + 1.f) * sqrt_d);
        }
        // mixing vectors: a thread owns (query head, latent channel) and all nsp keys - its W_comb row is read once per frame instead
        // of once per key (each output still sums c ascending)
        for (int o = tid; o < hq * Dl; o += nthr) {
            const int q = o / Dl, e = o - q * Dl, hk = q / g;
            const float* w = wc + (size_t)e * WP + q * D;
            for (int j0 = 0; j0 < nsp; j0 += 4) {          // four keys at a time (registers: the thread also holds its 64-float query row)
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
                for (int c = 0; c < D; c += 4) {
                    const float4 wv = *reinterpret_cast<const float4*>(w + c);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (j0 + j < nsp) {
                            const float4 vv = *reinterpret_cast<const float4*>(vs + ((j0 + j) * h + hk) * dp + c);
                            acc[j] = fmaf(wv.x, vv.x, acc[j]); acc[j] = fmaf(wv.y, vv.y, acc[j]); acc[j] = fmaf(wv.z, vv.z, acc[j]); acc[j] = fmaf(wv.w, vv.w, acc[j]);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) if (j0 + j < nsp) ys[((j0 + j) * hq + q) * Dl + e] = acc[j];
            }
        }
        __syncthreads();
        if (has_pair) {
            float s[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) s[j] = 0.f;
#pragma unroll
            for (int c = 0; c < D; c += 4) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (j < nsp) {
                        const float4 kv = *reinterpret_cast<const float4*>(ks + (j * h + my_hk) * dp + c);
                        s[j] = fmaf(qreg[c], kv.x, s[j]); s[j] = fmaf(qreg[c + 1], kv.y, s[j]); s[j] = fmaf(qreg[c + 2], kv.z, s[j]); s[j] = fmaf(qreg[c + 3], kv.w, s[j]);
                    }
                }
            }
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 8; ++j) if (j < nsp) { s[j] *= a.scale; mx = fmaxf(mx, s[j]); }
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) if (j < nsp) { s[j] = expf(s[j] - mx); sum += s[j]; }
            const float gate = my_gate / sum;
#pragma unroll
            for (int j = 0; j < 8; ++j) if (j < nsp) cf[my_i * (JH + 1) + j * hq + my_q] = s[j] * gate;
        }
        __syncthreads();
        // four latent channels per thread: one broadcast coefficient + one 16-byte read per 4 FMAs (each output still sums jh ascending)
        float* outb = a.pred + (long long)b * N * Dl;
        for (int o4 = tid; o4 < N * (Dl / 4); o4 += nthr) {
            const int e4 = o4 / N, i = o4 - e4 * N, e = e4 * 4;          // lanes = consecutive queries: ys is a broadcast read, cf rows are 33 floats apart
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int jh = 0; jh < JH; ++jh) {
                const float cc = cf[i * (JH + 1) + jh];
                const float4 y = *reinterpret_cast<const float4*>(ys + jh * Dl + e);
                acc.x = fmaf(cc, y.x, acc.x); acc.y = fmaf(cc, y.y, acc.y); acc.z = fmaf(cc, y.z, acc.z); acc.w = fmaf(cc, y.w, acc.w);
            }
            *reinterpret_cast<float4*>(outb + (long long)i * Dl + e) = acc;
        }
    }
}

}  // namespace

int d4_l2s_fused_supported(const L2sArgs& a) {
    const int NQ = a.nsp * a.g;
    return (a.Dl == 16 || a.Dl == 32 || a.Dl == 64) && a.N <= L2S_MAXN && NQ <= L2S_MAXQ && a.d % 4 == 0 && a.d <= 128 && a.h <= 32 &&
           (reinterpret_cast<uintptr_t>(a.w_k) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.w_v) & 15) == 0;
}

template <int DL, bool MMA, int NQC>
static int launch_l2s_q(const L2sArgs& a, cudaStream_t s) {
    const int nwarps = a.h < 8 ? a.h : 8;
    const size_t smem = sizeof(float) * ((size_t)L2S_MAXN * (DL + 4) + (size_t)nwarps * (L2S_MAXN * NQC + NQC * DL + NQC * a.d + a.d * (DL + 4)));
    static size_t configured = 0;
    if (smem > configured) {
        D4_CUDA_OK(cudaFuncSetAttribute(l2s_fused_kernel<DL, MMA, NQC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    l2s_fused_kernel<DL, MMA, NQC><<<a.B, nwarps * 32, smem, s>>>(a);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int DL, bool MMA>
static int launch_l2s(const L2sArgs& a, cudaStream_t s) {      // 4-query tiles keep two CTAs per SM at the default 4 spatial tokens
    return (a.nsp * a.g <= 4) ? launch_l2s_q<DL, MMA, 4>(a, s) : launch_l2s_q<DL, MMA, 8>(a, s);
}

int d4_l2s_fused(const L2sArgs& a, cudaStream_t s) {
    if (a.B <= 0) return 0;
    if (!d4_l2s_fused_supported(a)) return d4_fail("l2s_fused: unsupported shape");
    const bool mma = a.allow_tensor && (a.d % 8 == 0);
    if (a.Dl == 16) return mma ? launch_l2s<16, true>(a, s) : launch_l2s<16, false>(a, s);
    if (a.Dl == 32) return mma ? launch_l2s<32, true>(a, s) : launch_l2s<32, false>(a, s);
    return mma ? launch_l2s<64, true>(a, s) : launch_l2s<64, false>(a, s);
}

static int g_lp_version = [] { const char* v = getenv("D4_LP_V"); return v ? atoi(v) : 2; }();          // 1: the per-frame kernel at any batch
void d4_lp_fused_debug(int version) { g_lp_version = version; }

int d4_lp_fused_supported(const LpArgs& a) {
    return a.nsp <= 8 && a.nsp * a.hq <= LP_MAXJH && a.d % 4 == 0 && a.N * a.hq <= 4096 && (a.hq % a.h) == 0 &&
           (reinterpret_cast<uintptr_t>(a.w_comb) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.q) & 15) == 0;
}

int d4_lp_fused(const LpArgs& a, cudaStream_t s) {
    if (a.B <= 0) return 0;
    if (!d4_lp_fused_supported(a)) return d4_fail("lp_fused: unsupported shape");
    const int JH = a.nsp * a.hq;
    const size_t smem = sizeof(float) * ((size_t)2 * a.nsp * a.h * (a.d + 4) + (size_t)LP_MAXJH * a.Dl + (size_t)a.N * (JH + 1));
    static size_t configured = 0;
    if (smem > configured) {
        if (smem > 200 * 1024) return d4_fail("lp_fused: %zu bytes of shared memory needed", smem);
        D4_CUDA_OK(cudaFuncSetAttribute(lp_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    // large batches, head dim 64: the persistent kernel (queries in registers, W_comb in shared memory); D4_LP_V=1 / d4_debug_set("lp_fused", 1) keeps this one
    static int num_sms = 0;
    if (!num_sms) { int dev = 0; D4_CUDA_OK(cudaGetDevice(&dev)); D4_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev)); }
    const int lp_v = g_lp_version;
    const size_t smem_p = smem + sizeof(float) * (size_t)a.Dl * (a.hq * a.d + 4);
    if (lp_v != 1 && a.d == 64 && a.N * a.hq <= LPP_THREADS && a.B > num_sms && a.Dl % 4 == 0 && (reinterpret_cast<uintptr_t>(a.pred) & 15) == 0 && smem_p <= 220 * 1024) {
        static size_t configured_p = 0;
        if (smem_p > configured_p) {
            D4_CUDA_OK(cudaFuncSetAttribute(lp_fused_persist_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));
            configured_p = smem_p;
        }
        lp_fused_persist_kernel<64><<<num_sms, LPP_THREADS, smem_p, s>>>(a);
        D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError());
        return 0;
    }
    lp_fused_kernel<<<a.B, 256, smem, s>>>(a);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError());
    return 0;
}
