// K1 variant 1: time-decode attention with the KV stream staged by the bulk-copy engine.
//
// Persistent kernel, one CTA per SM, 16 warps; every warp owns a private ring of shared-memory stages fed by
// cp.async.bulk (UBLKCP) completing on warp-private mbarriers, and walks its (token, kv-head) streams with the ring
// always NS tiles ahead — across stream boundaries too, so the HBM pipe never drains between streams.
// A tile is 16 cached keys (or values) = 16*d*4 contiguous bytes of the (M, h, Tmax, d) cache.
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

constexpr int TB_TK = 16;            // cached key (or value) rows per full tile
constexpr int TB_RING = 48;          // ring capacity per warp in cached rows: three full tiles in flight / in use
constexpr int TB_MAXNS = 16;         // short contexts slice the same ring into up to 16 smaller stages
constexpr int TB_MAXT = 256;         // longest context (scores of one stream live in shared memory)

template <int D, int G>
struct TbWarpSmem {
    float tile[TB_RING * D];
    float s[G][TB_MAXT];             // scores, then unnormalised probabilities, of the current stream
    float q[G * D];
    float vnew[D];
    uint64_t bar[TB_MAXNS];
};

// the new token's inputs for one (token, kv head) stream, in rotary-pair layout (lane p owns elements p and p + D/2)
template <int PPL, int G>
struct TbRaw {
    float k1[PPL], k2[PPL], v1[PPL], v2[PPL], r1[PPL], r2[PPL], g1[PPL], g2[PPL];
    float q1[G][PPL], q2[G][PPL];
    float mix, gate[G];
    int m, hk;
};

template <int D, int G, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) time_attn_bulk_kernel(TimeAttnArgs a) {
    constexpr int HALF = D / 2;
    constexpr int PPL = (HALF + 31) / 32;
    constexpr int CH = D / 8;             // float4 chunks per half row (two lanes share one cached key in the score pass)
    constexpr int LPK = D / 4;            // lanes that cover one cached value row with float4 columns
    constexpr int KG = 32 / LPK;          // value rows processed per AV step
    static_assert(D % 8 == 0 && LPK <= 32 && (32 % LPK) == 0, "head dim must be 16, 32, 64 or 128");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    TbWarpSmem<D, G>& sm = reinterpret_cast<TbWarpSmem<D, G>*>(smem_raw)[warp];

    const int total = a.M * a.hkv;
    const int wstride = (int)gridDim.x * WARPS;
    const int first = (int)blockIdx.x * WARPS + warp;
    const int t = a.t;
    // a tile is TK cached rows; contexts shorter than a full tile use one small tile per stream and a deeper ring, so the
    // copies in flight still cover several streams ahead
    const int TK = min(max(t, 1), TB_TK);         // (nvcc 12.9 folds the equivalent `t >= 16 ? 16 : max(t, 1)` to max(t, 1): keep this form)
    const int nt = (t + TK - 1) / TK;             // tiles per K (and per V) stream
    const int NS = min(TB_MAXNS, TB_RING / TK);
    const int n_items = (first < total) ? (total - first + wstride - 1) / wstride : 0;
    const int n_tiles = n_items * 2 * nt;

    if (lane == 0) {
        for (int s = 0; s < TB_MAXNS; ++s) mbar_init(&sm.bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    // producer side (lane 0): the next tile to request, as (stream, K/V phase, tile) plus its ring slot
    int p_item = first, p_ph = 0, p_ti = 0, p_slot = 0, p_idx = 0;
    // stream -> (token, kv head): the launch may cover a subset of the token rows (a.tmap), e.g. the spatial tokens of every frame
    auto stream_row = [&](int item) { const int mc = item / a.hkv; return (long long)a.tmap(mc) * a.hkv + (item - mc * a.hkv); };
    auto issue = [&]() {
        const float* base = (p_ph == 0 ? a.kcache : a.vcache) + (stream_row(p_item) * a.Tmax + (long long)p_ti * TK) * D;
        const uint32_t bytes = (uint32_t)min(TK, t - p_ti * TK) * D * 4;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&sm.bar[p_slot], bytes);
        bulk_g2s(sm.tile + p_slot * TK * D, base, bytes, &sm.bar[p_slot]);
        if (++p_ti == nt) { p_ti = 0; if (++p_ph == 2) { p_ph = 0; p_item += wstride; } }
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
    const bool clamp = a.softclamp > 0.f;
    const float inv_clamp = clamp ? 1.f / a.softclamp : 0.f;

    using Raw = TbRaw<PPL, G>;
    auto load_raw = [&](int item, Raw& r) {
        const int mc = item / a.hkv, hk = item - mc * a.hkv;
        const int m = (int)a.tmap(mc);
        r.m = m; r.hk = hk;
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
#pragma unroll 1
    for (int li = 0; li < n_items; ++li) {
        const int item = first + li * wstride;
        // the next stream's inputs are requested now and consumed one iteration later: their latency hides under this stream
        Raw nxt;
        if (li + 1 < n_items) load_raw(item + wstride, nxt);
        const int m = cur.m, hk = cur.hk;

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
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) { ss += __shfl_xor_sync(D4_FULL, ss, off); vs += __shfl_xor_sync(D4_FULL, vs, off); }
        const float rk = 1.f / fmaxf(sqrtf(ss), D4_L2_EPS);
        const float rv = 1.f / fmaxf(sqrtf(vs), D4_L2_EPS);
#pragma unroll
        for (int e = 0; e < PPL; ++e) {
            const int p = lane + 32 * e;
            const float n1 = (cur.k1[e] * rk) * ((cur.g1[e] + 1.f) * sqrt_d);
            const float n2 = (cur.k2[e] * rk) * ((cur.g2[e] + 1.f) * sqrt_d);
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
            if (clamp) s = tanhf(s * inv_clamp) * a.softclamp;
            self_s[gi] = s;
        }
        __syncwarp();

        // ---- scores over the cached keys: two lanes per key (half a row each), rotated float4 chunk order keeps the
        //      linear tile conflict free; the scores of the whole stream are parked in shared memory
        {
            const int key = lane >> 1, half = lane & 1;
#pragma unroll 1
            for (int ti = 0; ti < nt; ++ti) {
                const int nk = min(TK, t - ti * TK);
                mbar_wait(&sm.bar[c_slot], c_phase);
                const float* tile = sm.tile + c_slot * TK * D;
                float acc[G];
#pragma unroll
                for (int gi = 0; gi < G; ++gi) acc[gi] = 0.f;
                if (key < nk) {
                    const float* kr = tile + key * D + half * HALF;
#pragma unroll
                    for (int c = 0; c < CH; ++c) {
                        const int cc = ((c + lane) % CH) * 4;
                        const float4 kv = *reinterpret_cast<const float4*>(kr + cc);
#pragma unroll
                        for (int gi = 0; gi < G; ++gi) {
                            const float4 qv = *reinterpret_cast<const float4*>(sm.q + gi * D + half * HALF + cc);
                            acc[gi] = fmaf(qv.x, kv.x, acc[gi]); acc[gi] = fmaf(qv.y, kv.y, acc[gi]);
                            acc[gi] = fmaf(qv.z, kv.z, acc[gi]); acc[gi] = fmaf(qv.w, kv.w, acc[gi]);
                        }
                    }
                }
#pragma unroll
                for (int gi = 0; gi < G; ++gi) {
                    float sv = (acc[gi] + __shfl_xor_sync(D4_FULL, acc[gi], 1)) * a.scale;
                    if (clamp) sv = tanhf(sv * inv_clamp) * a.softclamp;
                    if (half == 0 && key < nk) sm.s[gi][ti * TK + key] = sv;
                }
                __syncwarp();
                if (lane == 0 && p_idx < n_tiles) issue();
                if (++c_slot == NS) { c_slot = 0; c_phase ^= 1; }
            }
        }

        // ---- softmax over cached keys + self (normalisation is folded into the output)
        float es[G], inv[G];
#pragma unroll
        for (int gi = 0; gi < G; ++gi) {
            float mx = self_s[gi];
            for (int j = lane; j < t; j += 32) mx = fmaxf(mx, sm.s[gi][j]);
            mx = warp_max(mx);
            float sum = 0.f;
            for (int j = lane; j < t; j += 32) { const float e = expf(sm.s[gi][j] - mx); sm.s[gi][j] = e; sum += e; }
            sum = warp_sum(sum);
            es[gi] = expf(self_s[gi] - mx);
            inv[gi] = 1.f / (sum + es[gi]);
        }
        __syncwarp();

        // ---- AV over the cached values: KG value rows per step, lane = (row group, float4 column), probabilities from smem
        const int kg = lane / LPK, c4 = (lane % LPK) * 4;
        float4 o[G];
#pragma unroll
        for (int gi = 0; gi < G; ++gi) o[gi] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
        for (int ti = 0; ti < nt; ++ti) {
            const int nk = min(TK, t - ti * TK);
            mbar_wait(&sm.bar[c_slot], c_phase);
            const float* tile = sm.tile + c_slot * TK * D;
#pragma unroll 4
            for (int j = kg; j < nk; j += KG) {
                const float4 vv = *reinterpret_cast<const float4*>(tile + j * D + c4);
#pragma unroll
                for (int gi = 0; gi < G; ++gi) {
                    const float pj = sm.s[gi][ti * TK + j];
                    o[gi].x = fmaf(pj, vv.x, o[gi].x); o[gi].y = fmaf(pj, vv.y, o[gi].y);
                    o[gi].z = fmaf(pj, vv.z, o[gi].z); o[gi].w = fmaf(pj, vv.w, o[gi].w);
                }
            }
            __syncwarp();
            if (lane == 0 && p_idx < n_tiles) issue();
            if (++c_slot == NS) { c_slot = 0; c_phase ^= 1; }
        }
#pragma unroll
        for (int gi = 0; gi < G; ++gi) {
#pragma unroll
            for (int off = LPK; off < 32; off <<= 1) {
                o[gi].x += __shfl_xor_sync(D4_FULL, o[gi].x, off); o[gi].y += __shfl_xor_sync(D4_FULL, o[gi].y, off);
                o[gi].z += __shfl_xor_sync(D4_FULL, o[gi].z, off); o[gi].w += __shfl_xor_sync(D4_FULL, o[gi].w, off);
            }
        }

        // ---- epilogue (float4-column layout): + self, softmax normalisation, belief projection on the new value, head gate
        const float4 vn = *reinterpret_cast<const float4*>(sm.vnew + c4);
        const float4 vh = make_float4(vn.x * rv, vn.y * rv, vn.z * rv, vn.w * rv);
#pragma unroll
        for (int gi = 0; gi < G; ++gi) {
            const int hq = hk * G + gi;
            o[gi].x = fmaf(es[gi], vn.x, o[gi].x) * inv[gi]; o[gi].y = fmaf(es[gi], vn.y, o[gi].y) * inv[gi];
            o[gi].z = fmaf(es[gi], vn.z, o[gi].z) * inv[gi]; o[gi].w = fmaf(es[gi], vn.w, o[gi].w) * inv[gi];
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
            float* kd = a.kcache + (((long long)m * a.hkv + hk) * a.Tmax + t) * D;
            float* vd = a.vcache + (((long long)m * a.hkv + hk) * a.Tmax + t) * D;
#pragma unroll
            for (int e = 0; e < PPL; ++e) {
                const int p = lane + 32 * e;
                if (p < HALF) { kd[p] = k1[e]; kd[p + HALF] = k2[e]; vd[p] = v1[e]; vd[p + HALF] = v2[e]; }
            }
        }
        __syncwarp();      // sm.q / sm.vnew / sm.s are rewritten by the next stream
        cur = nxt;
    }
}

template <int D, int G>
int launch_bulk(const TimeAttnArgs& a, cudaStream_t s) {
    // as many warps per SM as the per-warp ring leaves room for (16 at the config-4 shape)
    constexpr int WARPS = (sizeof(TbWarpSmem<D, G>) * 16 <= 227 * 1024) ? 16 : 12;
    const size_t smem = sizeof(TbWarpSmem<D, G>) * WARPS;
    static bool configured = false;
    static int num_sms = 0;
    if (!configured) {
        D4_CUDA_OK(cudaFuncSetAttribute(time_attn_bulk_kernel<D, G, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int dev = 0;
        D4_CUDA_OK(cudaGetDevice(&dev));
        D4_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        configured = true;
    }
    const long long items = (long long)a.M * a.hkv;
    if (items >= (1ll << 31) / 2) return d4_fail("time_attn(bulk): %lld streams exceed the 32-bit stream index", items);
    const long long ctas_needed = (items + WARPS - 1) / WARPS;
    const unsigned grid = (unsigned)std::min<long long>(ctas_needed, num_sms);
    time_attn_bulk_kernel<D, G, WARPS><<<grid, WARPS * 32, smem, s>>>(a);
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
