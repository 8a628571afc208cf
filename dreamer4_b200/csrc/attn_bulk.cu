// K1 variant 1: time-decode attention with the KV stream staged by the bulk-copy engine.
//
// Persistent kernel, one CTA per SM, 8 warps; every warp owns a private ring of NS shared-memory stages fed by
// cp.async.bulk (UBLKCP) completing on warp-private mbarriers, and walks its (token, kv-head) streams with the ring
// always NS tiles ahead — across stream boundaries too, so the HBM pipe never drains between streams.
// Each tile is 32 cached keys (or values) = 32*d*4 contiguous bytes of the (M, h, Tmax, d) cache.
// The next stream's token inputs are prefetched into registers while the current stream is processed, so the only
// exposed latency is the bulk-copy ring itself.
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
constexpr int TB_NS = 3;             // ring depth in full 32-key tiles: the ring is TB_NS * 32 key rows per warp
constexpr int TB_MAXNS = 16;         // short contexts slice the same shared memory into up to 16 smaller stages
constexpr int TB_MAXTILES = 8;

template <int D, int G>
struct TbWarpSmem {
    float tile[TB_NS * 32 * D];
    float q[G * D];
    float vnew[D];
    float p[G][32];
    uint64_t bar[TB_MAXNS];
};

// the new token's inputs for one (token, kv head) stream, in rotary-pair layout (lane p owns elements p and p + D/2)
template <int PPL, int G>
struct TbRaw {
    float k1[PPL], k2[PPL], v1[PPL], v2[PPL], r1[PPL], r2[PPL], g1[PPL], g2[PPL];
    float q1[G][PPL], q2[G][PPL];
    float mix, gate[G];
};

template <int D, int G>
__global__ void __launch_bounds__(TB_WARPS * 32, 1) time_attn_bulk_kernel(TimeAttnArgs a) {
    constexpr int HALF = D / 2;
    constexpr int PPL = (HALF + 31) / 32;
    constexpr int C4 = D / 4;
    constexpr int LPK = D / 4;            // lanes that cover one cached value row with float4 columns
    constexpr int KG = 32 / LPK;          // value rows processed per AV step
    static_assert(D % 4 == 0 && LPK <= 32 && (32 % LPK) == 0, "head dim must be 16, 32, 64 or 128");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    TbWarpSmem<D, G>& sm = reinterpret_cast<TbWarpSmem<D, G>*>(smem_raw)[warp];

    const long long total = (long long)a.M * a.hkv;
    const long long wstride = (long long)gridDim.x * TB_WARPS;
    const long long first = (long long)blockIdx.x * TB_WARPS + warp;
    const int t = a.t;
    // a tile is TK cached key (or value) rows; contexts shorter than 32 use one small tile per stream and a deeper ring,
    // so the copies in flight still cover several streams ahead
    const int TK = min(32, max(t, 1));
    const int nt = (t + TK - 1) / TK;             // tiles per K (and per V) stream
    const int NS = min(TB_MAXNS, (TB_NS * 32) / TK);
    const long long n_items = (first < total) ? (total - first + wstride - 1) / wstride : 0;
    const long long n_tiles = n_items * 2 * nt;

    if (lane == 0) {
        for (int s = 0; s < TB_MAXNS; ++s) mbar_init(&sm.bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    // producer side (lane 0): the next tile to request, as (stream, K/V phase, tile) plus its ring slot
    long long p_li = 0, p_idx = 0;
    int p_r = 0, p_slot = 0;
    auto issue = [&]() {
        const int ph = p_r / nt, ti = p_r - ph * nt;
        const long long item = first + p_li * wstride;
        const float* base = (ph == 0 ? a.kcache : a.vcache) + item * (long long)a.Tmax * D + (long long)ti * TK * D;
        const int nk = min(TK, t - ti * TK);
        const uint32_t bytes = (uint32_t)nk * D * 4;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&sm.bar[p_slot], bytes);
        bulk_g2s(sm.tile + p_slot * TK * D, base, bytes, &sm.bar[p_slot]);
        if (++p_r == 2 * nt) { p_r = 0; ++p_li; }
        if (++p_slot == NS) p_slot = 0;
        ++p_idx;
    };
    if (lane == 0)
        while (p_idx < NS && p_idx < n_tiles) issue();
    int c_slot = 0; uint32_t c_phase = 0;          // consumer ring position

    // rotary angles of position t: the same for every stream of this launch
    float cs[PPL], sn[PPL];
#pragma unroll
    for (int e = 0; e < PPL; ++e) {
        const int p = lane + 32 * e;
        cs[e] = 1.f; sn[e] = 0.f;
        if (p < HALF) sincosf((float)t * a.inv_freq[p], &sn[e], &cs[e]);
    }
    const float sqrt_d = sqrtf((float)D);

    using Raw = TbRaw<PPL, G>;
    auto load_raw = [&](long long item, Raw& r) {
        const int m = (int)(item / a.hkv), hk = (int)(item % a.hkv);
        const float* row = a.qkvgm + (long long)m * a.ld;
        const float* v0r = a.v0 + (long long)m * a.ldv0 + hk * D;
        r.mix = row[a.off_m + hk];
#pragma unroll
        for (int gi = 0; gi < G; ++gi) r.gate[gi] = row[a.off_g + hk * G + gi];
#pragma unroll
        for (int e = 0; e < PPL; ++e) {
            const int p = lane + 32 * e;
            const bool ok = p < HALF;
            r.k1[e] = ok ? row[a.off_k + hk * D + p] : 0.f; r.k2[e] = ok ? row[a.off_k + hk * D + p + HALF] : 0.f;
            r.v1[e] = ok ? row[a.off_v + hk * D + p] : 0.f; r.v2[e] = ok ? row[a.off_v + hk * D + p + HALF] : 0.f;
            r.r1[e] = ok ? v0r[p] : 0.f; r.r2[e] = ok ? v0r[p + HALF] : 0.f;
            r.g1[e] = ok ? a.k_gamma[hk * D + p] : 0.f; r.g2[e] = ok ? a.k_gamma[hk * D + p + HALF] : 0.f;
#pragma unroll
            for (int gi = 0; gi < G; ++gi) {
                r.q1[gi][e] = ok ? row[(hk * G + gi) * D + p] : 0.f;
                r.q2[gi][e] = ok ? row[(hk * G + gi) * D + p + HALF] : 0.f;
            }
        }
    };

    Raw cur;
    if (n_items > 0) load_raw(first, cur);
    for (long long li = 0; li < n_items; ++li) {
        const long long item = first + li * wstride;
        const int m = (int)(item / a.hkv), hk = (int)(item % a.hkv);
        // the next stream's inputs are requested now and consumed one iteration later: their latency hides under this stream
        Raw nxt;
        if (li + 1 < n_items) load_raw(item + wstride, nxt);

        // ---- prologue: value-residual lerp, key head-norm, rotary on q and k, the self score
        float k1[PPL], k2[PPL], v1[PPL], v2[PPL];
        const float mixw = 1.0f / (1.0f + expf(-cur.mix));
        float ss = 0.f, vs = 0.f;
#pragma unroll
        for (int e = 0; e < PPL; ++e) {
            v1[e] = lerp_(cur.v1[e], cur.r1[e], mixw);
            v2[e] = lerp_(cur.v2[e], cur.r2[e], mixw);
            ss += cur.k1[e] * cur.k1[e] + cur.k2[e] * cur.k2[e];
            vs += v1[e] * v1[e] + v2[e] * v2[e];
        }
        const float kden = fmaxf(sqrtf(warp_sum(ss)), D4_L2_EPS);
        const float vden = fmaxf(sqrtf(warp_sum(vs)), D4_L2_EPS);
#pragma unroll
        for (int e = 0; e < PPL; ++e) {
            const int p = lane + 32 * e;
            const float n1 = (cur.k1[e] / kden) * ((cur.g1[e] + 1.f) * sqrt_d);
            const float n2 = (cur.k2[e] / kden) * ((cur.g2[e] + 1.f) * sqrt_d);
            k1[e] = n1 * cs[e] + (-n2) * sn[e];
            k2[e] = n2 * cs[e] + n1 * sn[e];
            if (p < HALF) { sm.vnew[p] = v1[e]; sm.vnew[p + HALF] = v2[e]; }
        }
        float self_s[G];
#pragma unroll
        for (int gi = 0; gi < G; ++gi) {
            float dot = 0.f;
#pragma unroll
            for (int e = 0; e < PPL; ++e) {
                const int p = lane + 32 * e;
                const float r1 = cur.q1[gi][e] * cs[e] + (-cur.q2[gi][e]) * sn[e];
                const float r2 = cur.q2[gi][e] * cs[e] + cur.q1[gi][e] * sn[e];
                if (p < HALF) { sm.q[gi * D + p] = r1; sm.q[gi * D + p + HALF] = r2; }
                dot += r1 * k1[e] + r2 * k2[e];
            }
            float s = warp_sum(dot) * a.scale;
            if (a.softclamp > 0.f) s = tanhf(s / a.softclamp) * a.softclamp;
            self_s[gi] = s;
        }
        __syncwarp();

        // ---- scores over the cached keys: lane = key, rotated float4 chunk order keeps the linear tile conflict free
        float sc[G][TB_MAXTILES];
#pragma unroll
        for (int ti = 0; ti < TB_MAXTILES; ++ti) {
#pragma unroll
            for (int gi = 0; gi < G; ++gi) sc[gi][ti] = -INFINITY;
            if (ti < nt) {
                const int nk = min(TK, t - ti * TK);
                mbar_wait(&sm.bar[c_slot], c_phase);
                const float* tile = sm.tile + c_slot * TK * D;
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
                if (lane == 0 && p_idx < n_tiles) issue();
                if (++c_slot == NS) { c_slot = 0; c_phase ^= 1; }
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

        // ---- AV over the cached values: KG value rows per step, lane = (row group, float4 column), probabilities broadcast from smem
        const int kg = lane / LPK, c4 = (lane % LPK) * 4;
        float4 o[G];
#pragma unroll
        for (int gi = 0; gi < G; ++gi) o[gi] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int ti = 0; ti < TB_MAXTILES; ++ti) {
            if (ti < nt) {
                const int nk = min(TK, t - ti * TK);
#pragma unroll
                for (int gi = 0; gi < G; ++gi) sm.p[gi][lane] = sc[gi][ti];       // zero beyond nk
                mbar_wait(&sm.bar[c_slot], c_phase);
                __syncwarp();
                const float* tile = sm.tile + c_slot * TK * D;
#pragma unroll 4
                for (int j = kg; j < nk; j += KG) {
                    const float4 vv = *reinterpret_cast<const float4*>(tile + j * D + c4);
#pragma unroll
                    for (int gi = 0; gi < G; ++gi) {
                        const float pj = sm.p[gi][j];
                        o[gi].x = fmaf(pj, vv.x, o[gi].x); o[gi].y = fmaf(pj, vv.y, o[gi].y);
                        o[gi].z = fmaf(pj, vv.z, o[gi].z); o[gi].w = fmaf(pj, vv.w, o[gi].w);
                    }
                }
                __syncwarp();
                if (lane == 0 && p_idx < n_tiles) issue();
                if (++c_slot == NS) { c_slot = 0; c_phase ^= 1; }
            }
        }
#pragma unroll
        for (int gi = 0; gi < G; ++gi) {
#pragma unroll
            for (int off = LPK; off < 32; off <<= 1) {
                o[gi].x += __shfl_xor_sync(D4_FULL, o[gi].x, off); o[gi].y += __shfl_xor_sync(D4_FULL, o[gi].y, off);
                o[gi].z += __shfl_xor_sync(D4_FULL, o[gi].z, off); o[gi].w += __shfl_xor_sync(D4_FULL, o[gi].w, off);
            }
        }

        // ---- epilogue (float4-column layout): + self, belief projection on the new value, head gate, store
        const float4 vn = *reinterpret_cast<const float4*>(sm.vnew + c4);
        const float4 vh = make_float4(vn.x / vden, vn.y / vden, vn.z / vden, vn.w / vden);
#pragma unroll
        for (int gi = 0; gi < G; ++gi) {
            const int hq = hk * G + gi;
            o[gi].x = fmaf(pself[gi], vn.x, o[gi].x); o[gi].y = fmaf(pself[gi], vn.y, o[gi].y);
            o[gi].z = fmaf(pself[gi], vn.z, o[gi].z); o[gi].w = fmaf(pself[gi], vn.w, o[gi].w);
            float dot = o[gi].x * vh.x + o[gi].y * vh.y + o[gi].z * vh.z + o[gi].w * vh.w;
#pragma unroll
            for (int off = LPK / 2; off > 0; off >>= 1) dot += __shfl_xor_sync(D4_FULL, dot, off);
            const float gate = 1.0f / (1.0f + expf(-cur.gate[gi]));
            if (kg == 0) {
                float* op = a.out + (long long)m * a.ldo + hq * D + c4;
                *reinterpret_cast<float4*>(op) = make_float4((o[gi].x - dot * vh.x) * gate, (o[gi].y - dot * vh.y) * gate,
                                                             (o[gi].z - dot * vh.z) * gate, (o[gi].w - dot * vh.w) * gate);
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
        __syncwarp();      // sm.q / sm.vnew / sm.p are rewritten by the next stream
        cur = nxt;
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
    if ((reinterpret_cast<uintptr_t>(a.kcache) & 15) || (reinterpret_cast<uintptr_t>(a.vcache) & 15))
        return d4_fail("time_attn(bulk): KV cache must be 16-byte aligned");
    if ((reinterpret_cast<uintptr_t>(a.out) & 15) || (a.ldo & 3))
        return d4_fail("time_attn(bulk): output rows must be 16-byte aligned");
#define D4_TB_CASE(DD, GG) if (a.d == DD && a.g == GG) return launch_bulk<DD, GG>(a, s);
    D4_TB_CASE(64, 1) D4_TB_CASE(64, 2) D4_TB_CASE(32, 1) D4_TB_CASE(32, 2) D4_TB_CASE(16, 1) D4_TB_CASE(16, 2)
#undef D4_TB_CASE
    return d4_fail("time_attn(bulk): (dim_head=%d, query groups=%d) has no kernel instantiation", a.d, a.g);
}
