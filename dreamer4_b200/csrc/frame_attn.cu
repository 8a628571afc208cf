// Attention within one frame when the frame has more tokens than small_attn_kernel's 64 keys: the video tokenizer's encoder /
// decoder transformers see patches + latent tokens per frame (128 at 256 x 256 / patch 32, reference dreamer4.py:4360, 3655).
//
// Same arithmetic as small_attn_kernel in attn.cu (reference dreamer4.py:1968-2075 + naive_attend 1683-1756: key
// MultiHeadRMSNorm, softclamp, special-token mask 1769-1783, value-residual lerp 2005-2012, belief projection 2049-2054, head
// gates), restructured for a large key set: ONE CTA per (frame, kv head) stages that head's K and V once in shared memory,
// its warps then take the queries round-robin; a query's scores live in a per-warp shared row instead of registers.
//   a.mask_agent carries the NUMBER of special tokens here: queries i < nq - ns do not see keys j >= n - ns.
//
// This first version is exact-fp32 FMA (the `fp32` engine mode, and the shapes the tensor-core version below does not take);
// 4 MFLOP per (frame, head) at n = 128, d = 64.
#include "kernels.h"
#include <float.h>

namespace {

__device__ __forceinline__ float lerp_(float a, float b, float w) {      // torch.lerp as ATen evaluates it
    const float d = b - a;
    return (w < 0.5f) ? a + w * d : b - d * (1.f - w);
}

constexpr int FA_WARPS = 8;

__global__ void __launch_bounds__(FA_WARPS * 32) frame_attn_kernel(SmallAttnArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x / a.hkv, hk = blockIdx.x % a.hkv;
    const int d = a.d, n = a.n, kp = d + 4, d4 = d >> 2, ns = a.mask_agent;
    float* Ks = smem;
    float* Vs = Ks + (size_t)n * kp;
    float* Ps = Vs + (size_t)n * kp + (size_t)warp * n;          // this warp's probability row

    // stage K, V of this (frame, head): value-residual lerp applied on the way in
    for (int idx = threadIdx.x; idx < n * d4; idx += FA_WARPS * 32) {
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
    __syncthreads();
    // MultiHeadRMSNorm on keys: l2norm(k) * (gamma + 1) * sqrt(d)   (reference dreamer4.py:1663-1679)
    const float sqrt_d = sqrtf((float)d);
    for (int j = threadIdx.x; j < n; j += FA_WARPS * 32) {
        float* kr = Ks + j * kp;
        float ss = 0.f;
        for (int c = 0; c < d; c += 4) { const float4 t = *reinterpret_cast<const float4*>(kr + c); ss += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w; }
        const float den = fmaxf(sqrtf(ss), D4_L2_EPS);
        for (int c = 0; c < d; ++c) kr[c] = (kr[c] / den) * ((a.k_gamma[hk * d + c] + 1.f) * sqrt_d);
    }
    __syncthreads();

    for (int gi = 0; gi < a.g; ++gi) {
        const int hq = hk * a.g + gi;
        for (int i = warp; i < a.nq; i += FA_WARPS) {
            const float* qp = a.q + b * a.q_sb + i * a.q_si + (long long)hq * d;
            // scores of query i against every key: lane owns keys lane, lane + 32, ...
            float mx = -INFINITY;
            for (int j = lane; j < n; j += 32) {
                const float* kr = Ks + j * kp;
                float acc = 0.f;
                for (int c = 0; c < d; c += 4) {
                    const float4 qv = __ldg(reinterpret_cast<const float4*>(qp + c));
                    const float4 kv = *reinterpret_cast<const float4*>(kr + c);
                    acc = fmaf(qv.x, kv.x, acc); acc = fmaf(qv.y, kv.y, acc); acc = fmaf(qv.z, kv.z, acc); acc = fmaf(qv.w, kv.w, acc);
                }
                float s = acc * a.scale;
                if (a.softclamp > 0.f) s = tanhf(s / a.softclamp) * a.softclamp;
                if (ns > 0 && i < a.nq - ns && j >= n - ns) s = -FLT_MAX;
                Ps[j] = s;
                mx = fmaxf(mx, s);
            }
            mx = warp_max(mx);
            float sum = 0.f;
            for (int j = lane; j < n; j += 32) { const float p = expf(Ps[j] - mx); Ps[j] = p; sum += p; }
            const float inv = 1.f / warp_sum(sum);
            __syncwarp();

            // out[c] = sum_j p_j V[j][c]; lane owns c = lane + 32 * e
            float o[4] = {0.f, 0.f, 0.f, 0.f};
            for (int j = 0; j < n; ++j) {
                const float pj = Ps[j];
#pragma unroll
                for (int e = 0; e < 4; ++e) { const int c = lane + 32 * e; if (c < d) o[e] = fmaf(pj, Vs[j * kp + c], o[e]); }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] *= inv;
            if (a.belief) {   // out -= (out . vhat) vhat,  vhat = l2norm(v_i)  (self-attention only: key i is token i)
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
            __syncwarp();                                           // Ps is rewritten by this warp's next query
        }
    }
}

inline bool al16p(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

int d4_frame_attn_mma_ok(const SmallAttnArgs& a);
int d4_frame_attn_mma(const SmallAttnArgs& a, cudaStream_t s);

int d4_frame_attn(const SmallAttnArgs& a, cudaStream_t s) {
    if (a.nb <= 0 || a.nq <= 0) return 0;
    if (d4_frame_attn_mma_ok(a)) return d4_frame_attn_mma(a, s);          // tensor-core engine modes, head dim 64, <= 128 keys
    if (a.n < 1) return d4_fail("frame_attn: no keys");
    if (a.d % 4 != 0 || a.d > 128) return d4_fail("frame_attn: head dim %d unsupported", a.d);
    if (a.belief && a.nq != a.n) return d4_fail("frame_attn: belief projection needs nq == n");
    if (a.mask_agent < 0 || a.mask_agent > a.n) return d4_fail("frame_attn: %d special tokens of %d", a.mask_agent, a.n);
    if (((a.q_sb | a.q_si | a.k_sb | a.k_sj | a.v_sb | a.v_sj | a.v0_sb | a.v0_sj) & 3) || !al16p(a.q) || !al16p(a.k) || !al16p(a.v) || (a.v0 && !al16p(a.v0)))
        return d4_fail("frame_attn: q / k / v rows must be 16-byte aligned");
    const size_t smem = ((size_t)2 * a.n * (a.d + 4) + (size_t)FA_WARPS * a.n) * sizeof(float);
    if (smem > 227 * 1024) return d4_fail("frame_attn: %d tokens per frame x head dim %d need %zu bytes of shared memory (> 227 KB)", a.n, a.d, smem);
    static size_t configured = 0;
    if (smem > configured) {
        D4_CUDA_OK(cudaFuncSetAttribute(frame_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    frame_attn_kernel<<<(unsigned)((long long)a.nb * a.hkv), FA_WARPS * 32, smem, s>>>(a);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError());
    return 0;
}
