// K1 variant 1: time-decode attention with the KV stream staged by the bulk-copy engine.
//
// Persistent kernel, one CTA per SM, 8 warps; every warp owns a private ring of NS shared-memory stages fed by
// cp.async.bulk (UBLKCP) completing on warp-private mbarriers, and walks its (token, kv-head) streams with the ring
// always NS tiles ahead — across stream boundaries too, so the HBM pipe never drains between streams.
// Each tile is 32 cached keys (or values) = 32*d*4 contiguous bytes of the (M, h, Tmax, d) cache.
// Same arithmetic, in the same order, as variant 0 in attn.cu.
#include "kernels.h"
#include <float.h>
#include <algorithm>

namespace {

__device__ __forceinline__ float lerp_(float a, float b, float w) {
    const float d = b - a;
    return (w < 0.5f) ? a + w * d : b - d * (1.f - w);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (int spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1 << 24)) __trap();      // a lost copy would otherwise hang the GPU box
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int TB_WARPS = 8;
constexpr int TB_NS = 3;
constexpr int TB_MAXTILES = 8;

template <int D, int G>
struct TbWarpSmem {
    float tile[TB_NS][32 * D];
    float q[G * D];
    uint64_t bar[TB_NS];
    uint64_t pad_;
};

template <int D, int G>
__global__ void __launch_bounds__(TB_WARPS * 32, 1) time_attn_bulk_kernel(TimeAttnArgs a) {
    constexpr int HALF = D / 2;
    constexpr int PPL = (HALF + 31) / 32;
    constexpr int C4 = D / 4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    TbWarpSmem<D, G>& sm = reinterpret_cast<TbWarpSmem<D, G>*>(smem_raw)[warp];

    const long long total = (long long)a.M * a.hkv;
    const long long wstride = (long long)gridDim.x * TB_WARPS;
    const long long first = (long long)blockIdx.x * TB_WARPS + warp;
    const int t = a.t;
    const int nt = (t + 31) / 32;                 // tiles per K (and per V) stream; t >= 1 here
    const long long n_items = (first < total) ? (total - first + wstride - 1) / wstride : 0;
    const long long n_tiles = n_items * 2 * nt;

    if (lane == 0) {
        for (int s = 0; s < TB_NS; ++s) mbar_init(&sm.bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    // producer side (lane 0): tile index -> source address
    auto issue = [&](long long tidx) {
        const long long li = tidx / (2 * nt);
        const int r = (int)(tidx % (2 * nt));
        const int ph = r / nt, ti = r % nt;
        const long long item = first + li * wstride;
        const float* base = (ph == 0 ? a.kcache : a.vcache) + item * (long long)a.Tmax * D + (long long)ti * 32 * D;
        const int nk = min(32, t - ti * 32);
        const uint32_t bytes = (uint32_t)nk * D * 4;
        const int s = (int)(tidx % TB_NS);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&sm.bar[s], bytes);
        bulk_g2s(sm.tile[s], base, bytes, &sm.bar[s]);
    };
    if (lane == 0)
        for (long long p = 0; p < TB_NS && p < n_tiles; ++p) issue(p);

    const float sqrt_d = sqrtf((float)D);
    long long cons = 0;       // next tile to consume
    for (long long li = 0; li < n_items; ++li) {
        const long long item = first + li * wstride;
        const int m = (int)(item / a.hkv), hk = (int)(item % a.hkv);
        const float* row = a.qkvgm + (long long)m * a.ld;

        // ---- prologue (identical to variant 0)
        float k1[PPL], k2[PPL], v1[PPL], v2[PPL], cs[PPL], sn[PPL];
        const float mixw = 1.0f / (1.0f + expf(-row[a.off_m + hk]));
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
                sincosf((float)t * a.inv_freq[p], &sn[e], &cs[e]);
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

        // ---- scores over the cached keys
        float sc[G][TB_MAXTILES];
#pragma unroll
        for (int ti = 0; ti < TB_MAXTILES; ++ti) {
#pragma unroll
            for (int gi = 0; gi < G; ++gi) sc[gi][ti] = -INFINITY;
            if (ti < nt) {
                const int nk = min(32, t - ti * 32);
                const int s = (int)(cons % TB_NS);
                mbar_wait(&sm.bar[s], (uint32_t)((cons / TB_NS) & 1));
                const float* tile = sm.tile[s];
                if (lane < nk) {
                    float acc[G];
#pragma unroll
                    for (int gi = 0; gi < G; ++gi) acc[gi] = 0.f;
#pragma unroll
                    for (int c = 0; c < C4; ++c) {
                        const int cc = (c + lane) % C4;
                        const float4 kv = *reinterpret_cast<const float4*>(tile + lane * D + cc * 4);
#pragma unroll
                        for (int gi = 0; gi < G; ++gi) {
                            const float4 qv = *reinterpret_cast<const float4*>(sm.q + gi * D + cc * 4);
                            acc[gi] = fmaf(qv.x, kv.x, acc[gi]); acc[gi] = fmaf(qv.y, kv.y, acc[gi]);
                            acc[gi] = fmaf(qv.z, kv.z, acc[gi]); acc[gi] = fmaf(qv.w, kv.w, acc[gi]);
                        }
                    }
#pragma unroll
                    for (int gi = 0; gi < G; ++gi) {
                        float sv = acc[gi] * a.scale;
                        if (a.softclamp > 0.f) sv = tanhf(sv / a.softclamp) * a.softclamp;
                        sc[gi][ti] = sv;
                    }
                }
                __syncwarp();
                if (lane == 0 && cons + TB_NS < n_tiles) issue(cons + TB_NS);
                ++cons;
            }
        }

        // ---- softmax over cached keys + self
        float pself[G];
#pragma unroll
        for (int gi = 0; gi < G; ++gi) {
            float mx = self_s[gi];
#pragma unroll
            for (int ti = 0; ti < TB_MAXTILES; ++ti) mx = fmaxf(mx, sc[gi][ti]);
            mx = warp_max(mx);
            float sum = 0.f;
#pragma unroll
            for (int ti = 0; ti < TB_MAXTILES; ++ti) { const float e = (sc[gi][ti] == -INFINITY) ? 0.f : expf(sc[gi][ti] - mx); sc[gi][ti] = e; sum += e; }
            sum = warp_sum(sum);
            const float es = expf(self_s[gi] - mx);
            const float inv = 1.f / (sum + es);
#pragma unroll
            for (int ti = 0; ti < TB_MAXTILES; ++ti) sc[gi][ti] *= inv;
            pself[gi] = es * inv;
        }

        // ---- AV over the cached values
        float o1[G][PPL], o2[G][PPL];
#pragma unroll
        for (int gi = 0; gi < G; ++gi)
#pragma unroll
            for (int e = 0; e < PPL; ++e) { o1[gi][e] = 0.f; o2[gi][e] = 0.f; }
#pragma unroll
        for (int ti = 0; ti < TB_MAXTILES; ++ti) {
            if (ti < nt) {
                const int nk = min(32, t - ti * 32);
                const int s = (int)(cons % TB_NS);
                mbar_wait(&sm.bar[s], (uint32_t)((cons / TB_NS) & 1));
                const float* tile = sm.tile[s];
                for (int j = 0; j < nk; ++j) {
#pragma unroll
                    for (int gi = 0; gi < G; ++gi) {
                        const float pj = __shfl_sync(D4_FULL, sc[gi][ti], j);
#pragma unroll
                        for (int e = 0; e < PPL; ++e) {
                            const int p = lane + 32 * e;
                            if (p < HALF) {
                                o1[gi][e] = fmaf(pj, tile[j * D + p], o1[gi][e]);
                                o2[gi][e] = fmaf(pj, tile[j * D + p + HALF], o2[gi][e]);
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0 && cons + TB_NS < n_tiles) issue(cons + TB_NS);
                ++cons;
            }
        }

        // ---- epilogue
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
            const float gate = 1.0f / (1.0f + expf(-row[a.off_g + hq]));
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
            float* kd = a.kcache + (item * a.Tmax + t) * D;
            float* vd = a.vcache + (item * a.Tmax + t) * D;
#pragma unroll
            for (int e = 0; e < PPL; ++e) {
                const int p = lane + 32 * e;
                if (p < HALF) { kd[p] = k1[e]; kd[p + HALF] = k2[e]; vd[p] = v1[e]; vd[p + HALF] = v2[e]; }
            }
        }
        __syncwarp();      // sm.q is rewritten by the next stream's prologue
    }
}

template <int D, int G>
int launch_bulk(const TimeAttnArgs& a, cudaStream_t s) {
    const size_t smem = sizeof(TbWarpSmem<D, G>) * TB_WARPS;
    static bool configured = false;
    static int num_sms = 0;
    if (!configured) {
        D4_CUDA_OK(cudaFuncSetAttribute(time_attn_bulk_kernel<D, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int dev = 0;
        D4_CUDA_OK(cudaGetDevice(&dev));
        D4_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        configured = true;
    }
    const long long items = (long long)a.M * a.hkv;
    const long long ctas_needed = (items + TB_WARPS - 1) / TB_WARPS;
    const unsigned grid = (unsigned)std::min<long long>(ctas_needed, num_sms);
    time_attn_bulk_kernel<D, G><<<grid, TB_WARPS * 32, smem, s>>>(a);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace

int d4_time_attn_bulk(const TimeAttnArgs& a, cudaStream_t s) {
    if (a.t == 0) {          // nothing cached yet: no stream to stage, variant 0 handles the self-only case
        TimeAttnArgs b = a; b.variant = 0;
        return d4_time_attn(b, s);
    }
    if ((reinterpret_cast<uintptr_t>(a.kcache) & 15) || (reinterpret_cast<uintptr_t>(a.vcache) & 15))
        return d4_fail("time_attn(bulk): KV cache must be 16-byte aligned");
#define D4_TB_CASE(DD, GG) if (a.d == DD && a.g == GG) return launch_bulk<DD, GG>(a, s);
    D4_TB_CASE(64, 1) D4_TB_CASE(64, 2) D4_TB_CASE(32, 1) D4_TB_CASE(32, 2) D4_TB_CASE(16, 1) D4_TB_CASE(16, 2)
#undef D4_TB_CASE
    return d4_fail("time_attn(bulk): (dim_head=%d, query groups=%d) has no kernel instantiation", a.d, a.g);
}
