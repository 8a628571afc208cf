// tcgen05 TF32 GEMM, persistent warp-specialised version (the throughput path of the linear layers).
//
//   One CTA per SM walks output tiles (n fastest, so an A tile stays L2-hot across its N tiles).
//   warp 0   TMA producer      cp.async.bulk.tensor (SWIZZLE_128B) -> NS-stage shared-memory ring
//   warp 1   MMA issuer        tcgen05.mma.kind::tf32 128 x BN x 8 into one of TWO TMEM accumulators (BN fp32 columns each)
//   warps 2-5 epilogue         tcgen05.ld 32x32b -> row scale -> warp-private smem transpose -> bias / residual / GLU ->
//                              fully coalesced 128-byte row stores; overlaps the next tile's main loop
//   warps 6-9 A splitter       (tf32x3 only) hi/lo split of the landed A tile in shared memory
//   Barriers: full/empty per smem stage, split_done per stage (x3), tmem_full/tmem_empty per accumulator.
#include <cuda.h>
#include <string.h>
#include <algorithm>
#include "kernels.h"
#include "tc_ptx.cuh"

namespace {

constexpr int BM = 128, BK = 32, UMMA_K = 8;
constexpr int A_TILE = BM * BK * 4;           // 16 KB
constexpr int NUM_THREADS = 320;
constexpr int STG_LD = 33;                    // epilogue staging row pitch (floats)

struct __align__(64) TmaMaps2 { CUtensorMap a, w, wlo; };

using namespace d4tc;
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) { return mbar_try_wait_cta(bar, parity); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) { mbar_wait_cta(bar, parity); }
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) { return make_desc_sw128(smem_addr); }

struct EpiArgs2 {
    float* C; long long ldc; int M, N;
    const float* bias; const float* row_scale; const float* residual; long long ldr;
    int act; RowMap cmap;
    int a_grp, nkb, n_tiles_m, n_tiles_n;
};

template <int TERMS, int BN> struct Cfg {
    static constexpr int W_TILE = BN * BK * 4;
    static constexpr int STAGE = (TERMS == 3) ? 2 * A_TILE + 2 * W_TILE : A_TILE + W_TILE;
    static constexpr int NS = (192 * 1024) / STAGE;                    // 4 (x1,BN=256), 6 (x1,128), 2 (x3,256), 3 (x3,128)
    static constexpr int STG_BYTES = 4 * 32 * STG_LD * 4;
    static constexpr int SMEM = NS * STAGE + STG_BYTES + 512 + 1024;
    static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
};

template <int TERMS, int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_tc2_kernel(const __grid_constant__ TmaMaps2 maps, const EpiArgs2 e) {
    using K = Cfg<TERMS, BN>;
    constexpr int NS = K::NS;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* stg_all = reinterpret_cast<float*>(smem + NS * K::STAGE);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NS * K::STAGE + K::STG_BYTES);
    // bars: full[NS] | empty[NS] | split[NS] | tmem_full[2] | tmem_empty[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * NS + 4);
    auto bar = [&](int i) { return smem_u32(&bars[i]); };
    constexpr int B_FULL = 0, B_EMPTY = NS, B_SPLIT = 2 * NS, B_TFULL = 3 * NS, B_TEMPTY = 3 * NS + 2;
    constexpr int T_A = 0, T_ALO = A_TILE, T_W = (TERMS == 3) ? 2 * A_TILE : A_TILE, T_WLO = T_W + K::W_TILE;
    auto tile = [&](int stage, int off) { return smem + stage * K::STAGE + off; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = e.nkb;
    const int total_tiles = e.n_tiles_m * e.n_tiles_n;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.w) : "memory");
        if (TERMS == 3) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.wlo) : "memory");
        for (int s = 0; s < NS; ++s) { mbar_init(bar(B_FULL + s), 1); mbar_init(bar(B_EMPTY + s), 1); mbar_init(bar(B_SPLIT + s), 128); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar(B_TFULL + b), 1); mbar_init(bar(B_TEMPTY + b), 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    if (warp == 0) {
        // ================= TMA producer
        if (elect_one()) {
            uint32_t kc = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int m0 = (t / e.n_tiles_n) * BM, n0 = (t % e.n_tiles_n) * BN;
                for (int kb = 0; kb < nkb; ++kb, ++kc) {
                    const int s = kc % NS; const uint32_t ph = (kc / NS) & 1;
                    mbar_wait(bar(B_EMPTY + s), ph ^ 1);
                    mbar_expect_tx(bar(B_FULL + s), A_TILE + (TERMS == 3 ? 2 : 1) * K::W_TILE);
                    if (e.a_grp == 0) tma_load_2d(smem_u32(tile(s, T_A)), &maps.a, bar(B_FULL + s), kb * BK, m0);
                    else              tma_load_3d(smem_u32(tile(s, T_A)), &maps.a, bar(B_FULL + s), kb * BK, 0, m0 / e.a_grp);
                    tma_load_2d(smem_u32(tile(s, T_W)), &maps.w, bar(B_FULL + s), kb * BK, n0);
                    if (TERMS == 3) tma_load_2d(smem_u32(tile(s, T_WLO)), &maps.wlo, bar(B_FULL + s), kb * BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer
        if (elect_one()) {
            uint32_t kc = 0, ac = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++ac) {
                const int buf = ac & 1; const uint32_t aph = (ac >> 1) & 1;
                mbar_wait(bar(B_TEMPTY + buf), aph ^ 1);                 // epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_c = tmem_base + (uint32_t)(buf * BN);
                for (int kb = 0; kb < nkb; ++kb, ++kc) {
                    const int s = kc % NS; const uint32_t ph = (kc / NS) & 1;
                    mbar_wait(bar((TERMS == 3 ? B_SPLIT : B_FULL) + s), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t da = make_desc(smem_u32(tile(s, T_A))), dw = make_desc(smem_u32(tile(s, T_W)));
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t koff = (uint64_t)((k * UMMA_K * 4) >> 4);
                        const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
                        if (TERMS == 3) {
                            const uint64_t dalo = make_desc(smem_u32(tile(s, T_ALO))), dwlo = make_desc(smem_u32(tile(s, T_WLO)));
                            umma_tf32(tmem_c, dalo + koff, dw + koff, K::IDESC, acc);
                            umma_tf32(tmem_c, da + koff, dwlo + koff, K::IDESC, 1u);
                            umma_tf32(tmem_c, da + koff, dw + koff, K::IDESC, 1u);
                        } else {
                            umma_tf32(tmem_c, da + koff, dw + koff, K::IDESC, acc);
                        }
                    }
                    umma_commit(bar(B_EMPTY + s));
                }
                umma_commit(bar(B_TFULL + buf));
            }
        }
    } else if (warp < 6) {
        // ================= epilogue
        // Per 32-column chunk: residual rows are prefetched into registers (32 independent coalesced loads per lane) one
        // chunk ahead, the accumulator chunk is read with tcgen05.ld, transposed through a warp-private padded smem tile
        // and written back as fully coalesced 128-byte row segments.
        const int quarter = warp & 3;
        float* stg = stg_all + (warp - 2) * 32 * STG_LD;
        const bool glu = (e.act == D4_ACT_GLU_SILU || e.act == D4_ACT_GLU_GELU);
        const float* __restrict__ resid = e.residual;
        const float* __restrict__ biasp = e.bias;
        float* __restrict__ Cp = e.C;
        uint32_t ac = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++ac) {
            const int m0 = (t / e.n_tiles_n) * BM, n0 = (t % e.n_tiles_n) * BN;
            const int buf = ac & 1; const uint32_t aph = (ac >> 1) & 1;
            const int rbase = m0 + quarter * 32;
            const int mrow = rbase + lane;
            const float rs = (mrow < e.M && e.row_scale) ? e.row_scale[mrow] : 1.f;
            const int crow_lane = (mrow < e.M) ? (int)e.cmap(mrow) : -1;      // physical output row of TMEM lane `lane`
            const int nrows = min(32, e.M - rbase);                          // warp-uniform (may be <= 0)
            if (!glu) {
                const bool ident = (e.cmap.grp == 0);                        // output rows are the tile rows themselves
                const bool has_res = (resid != nullptr);
                float res[32];
                auto prefetch = [&](int c0, float (&dst)[32]) {
                    const int col = n0 + c0 + lane;
                    const bool ok = has_res && col < e.N;
                    if (ident) {
#pragma unroll
                        for (int r = 0; r < 32; ++r) dst[r] = (ok && r < nrows) ? __ldg(resid + (long long)(rbase + r) * e.ldr + col) : 0.f;
                    } else {
#pragma unroll
                        for (int r = 0; r < 32; ++r) {
                            const int crow = __shfl_sync(0xffffffffu, crow_lane, r);
                            dst[r] = (ok && crow >= 0) ? __ldg(resid + (long long)crow * e.ldr + col) : 0.f;
                        }
                    }
                };
                if (has_res) prefetch(0, res);
                else {
#pragma unroll
                    for (int r = 0; r < 32; ++r) res[r] = 0.f;
                }
                mbar_wait(bar(B_TFULL + buf), aph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_c = tmem_base + (uint32_t)(buf * BN) + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    const int nb = n0 + c0;
                    if (nb >= e.N) break;                               // warp-uniform
                    float v[32];
                    tmem_ld32(tmem_c + (uint32_t)c0, v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) stg[lane * STG_LD + j] = v[j] * rs;
                    __syncwarp();
                    float resn[32];
                    const bool more = has_res && (c0 + 32 < BN) && (nb + 32 < e.N);
                    if (more) prefetch(c0 + 32, resn);
                    const int col = nb + lane;
                    const bool cok = col < e.N;
                    const float bv = (cok && biasp) ? __ldg(biasp + col) : 0.f;
                    if (ident) {
                        float* cp = Cp + (long long)rbase * e.ldc + col;
#pragma unroll
                        for (int r = 0; r < 32; ++r)
                            if (r < nrows && cok) cp[(long long)r * e.ldc] = stg[r * STG_LD + lane] + bv + res[r];
                    } else {
#pragma unroll
                        for (int r = 0; r < 32; ++r) {
                            const int crow = __shfl_sync(0xffffffffu, crow_lane, r);
                            if (r < nrows && cok) Cp[(long long)crow * e.ldc + col] = stg[r * STG_LD + lane] + bv + res[r];
                        }
                    }
                    __syncwarp();
                    if (more) {
#pragma unroll
                        for (int r = 0; r < 32; ++r) res[r] = resn[r];
                    }
                }
            } else {
                mbar_wait(bar(B_TFULL + buf), aph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_c = tmem_base + (uint32_t)(buf * BN) + ((uint32_t)(quarter * 32) << 16);
                const int on = e.N >> 1;
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 64) {
                    const int nb = n0 + c0;
                    if (nb >= e.N) break;
                    // bias of the 64 interleaved input columns of this chunk: lane holds columns nb+lane and nb+32+lane
                    const float b_lo = (biasp && nb + lane < e.N) ? __ldg(biasp + nb + lane) : 0.f;
                    const float b_hi = (biasp && nb + 32 + lane < e.N) ? __ldg(biasp + nb + 32 + lane) : 0.f;
                    float v[32], w[32];
                    tmem_ld32(tmem_c + (uint32_t)c0, v);
                    tmem_ld32(tmem_c + (uint32_t)(c0 + 32), w);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float x0 = fmaf(v[2 * j], rs, __shfl_sync(0xffffffffu, b_lo, 2 * j));
                        const float g0 = fmaf(v[2 * j + 1], rs, __shfl_sync(0xffffffffu, b_lo, 2 * j + 1));
                        const float x1 = fmaf(w[2 * j], rs, __shfl_sync(0xffffffffu, b_hi, 2 * j));
                        const float g1 = fmaf(w[2 * j + 1], rs, __shfl_sync(0xffffffffu, b_hi, 2 * j + 1));
                        stg[lane * STG_LD + j] = x0 * ((e.act == D4_ACT_GLU_SILU) ? siluf_(g0) : geluf_(g0));
                        stg[lane * STG_LD + 16 + j] = x1 * ((e.act == D4_ACT_GLU_SILU) ? siluf_(g1) : geluf_(g1));
                    }
                    __syncwarp();
                    const int col = (nb >> 1) + lane;
                    const bool cok = col < on;
                    if (e.cmap.grp == 0) {
                        float* cp = Cp + (long long)rbase * e.ldc + col;
#pragma unroll
                        for (int r = 0; r < 32; ++r)
                            if (r < nrows && cok) cp[(long long)r * e.ldc] = stg[r * STG_LD + lane];
                    } else {
#pragma unroll
                        for (int r = 0; r < 32; ++r) {
                            const int crow = __shfl_sync(0xffffffffu, crow_lane, r);
                            if (r < nrows && cok) Cp[(long long)crow * e.ldc + col] = stg[r * STG_LD + lane];
                        }
                    }
                    __syncwarp();
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(bar(B_TEMPTY + buf));
        }
    } else if (TERMS == 3) {
        // ================= A splitter (tf32x3)
        const int et = threadIdx.x - 192;          // 0..127
        uint32_t kc = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            for (int kb = 0; kb < nkb; ++kb, ++kc) {
                const int s = kc % NS; const uint32_t ph = (kc / NS) & 1;
                mbar_wait(bar(B_FULL + s), ph);
                float4* a = reinterpret_cast<float4*>(tile(s, T_A));
                float4* alo = reinterpret_cast<float4*>(tile(s, T_ALO));
#pragma unroll
                for (int j = 0; j < A_TILE / 16 / 128; ++j) {
                    const int idx = et + 128 * j;
                    const float4 v = a[idx];
                    float4 hi, lo;
                    hi.x = tf32_rna(v.x); lo.x = tf32_rna(v.x - hi.x);
                    hi.y = tf32_rna(v.y); lo.y = tf32_rna(v.y - hi.y);
                    hi.z = tf32_rna(v.z); lo.z = tf32_rna(v.z - hi.z);
                    hi.w = tf32_rna(v.w); lo.w = tf32_rna(v.w - hi.w);
                    a[idx] = hi; alo[idx] = lo;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(bar(B_SPLIT + s));
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
    }
}

// ---------------------------------------------------------------- host side
int encode_2d(CUtensorMap* map, const float* base, long long rows, long long K, long long ld, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return d4_fail("cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return d4_fail("cuTensorMapEncodeTiled(2d rows=%lld K=%lld ld=%lld) failed: %d", rows, K, ld, (int)r);
    return 0;
}
int encode_3d(CUtensorMap* map, const float* base, long long M, long long K, long long ld, const RowMap& rm) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return d4_fail("cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rm.grp, (cuuint64_t)(M / rm.grp)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)rm.gstride * ld * 4};
    cuuint32_t box[3] = {BK, (cuuint32_t)rm.grp, (cuuint32_t)(BM / rm.grp)};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base + (long long)rm.goff * ld), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return d4_fail("cuTensorMapEncodeTiled(3d) failed: %d", (int)r);
    return 0;
}

int num_sms() {
    static int n = 0;
    if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
    return n;
}

template <int TERMS, int BN>
int launch2(const GemmArgs& g, cudaStream_t stream) {
    using K = Cfg<TERMS, BN>;
    TmaMaps2 maps; memset(&maps, 0, sizeof(maps));
    if (g.amap.grp == 0) { int rc = encode_2d(&maps.a, g.A, g.M, g.K, g.lda, BM); if (rc) return rc; }
    else { int rc = encode_3d(&maps.a, g.A, g.M, g.K, g.lda, g.amap); if (rc) return rc; }
    { int rc = encode_2d(&maps.w, g.W, g.N, g.K, g.ldw, BN); if (rc) return rc; }
    if (TERMS == 3) { int rc = encode_2d(&maps.wlo, g.W_lo, g.N, g.K, g.ldw, BN); if (rc) return rc; }
    EpiArgs2 e;
    e.C = g.C; e.ldc = g.ldc; e.M = g.M; e.N = g.N; e.bias = g.bias; e.row_scale = g.row_scale; e.residual = g.residual; e.ldr = g.ldr;
    e.act = g.act; e.cmap = g.cmap; e.a_grp = g.amap.grp; e.nkb = (g.K + BK - 1) / BK;
    e.n_tiles_m = (g.M + BM - 1) / BM; e.n_tiles_n = (g.N + BN - 1) / BN;
    static bool configured = false;
    if (!configured) {
        D4_CUDA_OK(cudaFuncSetAttribute(gemm_tc2_kernel<TERMS, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM));
        configured = true;
    }
    const long long tiles = (long long)e.n_tiles_m * e.n_tiles_n;
    const unsigned grid = (unsigned)std::min<long long>(tiles, num_sms());
    gemm_tc2_kernel<TERMS, BN><<<grid, NUM_THREADS, K::SMEM, stream>>>(maps, e);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace

// persistent kernel entry; bn = 128 or 256 (0 = choose by padding waste)
int d4_gemm_tc2(const GemmArgs& g, int terms, int bn, cudaStream_t stream) {
    if (g.rs_mode || g.ss_out) return d4_fail("gemm_tc2: sum-of-squares row statistics are only implemented by the CTA-pair kernel");
    if (bn == 0) {
        const double w128 = (double)((g.N + 127) / 128 * 128) / g.N, w256 = (double)((g.N + 255) / 256 * 256) / g.N;
        const long long tiles256 = (long long)((g.M + BM - 1) / BM) * ((g.N + 255) / 256);
        bn = (w256 <= w128 * 1.06 && tiles256 >= num_sms()) ? 256 : 128;
    }
    if (terms == 3) return bn == 256 ? launch2<3, 256>(g, stream) : launch2<3, 128>(g, stream);
    return bn == 256 ? launch2<1, 256>(g, stream) : launch2<1, 128>(g, stream);
}
