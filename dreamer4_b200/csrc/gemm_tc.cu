// tcgen05 TF32 GEMM for the linear layers of the imagination pass:  C = epi(rowscale * (A @ W^T) + bias) (+ residual)
//
//   A (M,K) fp32 activations, W (N,K) fp32 weights (nn.Linear layout) — both K-major, read as TF32 by the 5th-gen
//   tensor cores straight from the fp32 bits, fp32 accumulation in TMEM.
//   CTA tile 128 x 128, K step 32 floats (= one 128-byte swizzle atom row), UMMA 128x128x8 (kind::tf32), 4 per K step.
//   TMA (cp.async.bulk.tensor, SWIZZLE_128B) fills a 3-stage shared-memory ring; warp 0 = TMA producer, warp 1 = MMA
//   issuer + TMEM allocator, warps 2..5 = epilogue (tcgen05.ld 32x32b -> fused epilogue -> global stores).
//   One output tile per CTA, two CTAs resident per SM so one CTA's epilogue overlaps the other's main loop.
//
//   TERMS == 3 (tf32x3, fp32-accurate): a = a_hi + a_lo, w = w_hi + w_lo with hi = the top 19 bits; the product is
//   a_hi*w_hi + a_hi*w_lo + a_lo*w_hi (the dropped a_lo*w_lo term is ~2^-22 relative).  W is pre-split on the host side
//   once per weight change; A is split IN SHARED MEMORY by the (otherwise idle) epilogue warps right after the TMA lands,
//   so activations are read from HBM/L2 exactly once.
#include <cuda.h>
#include <string.h>
#include <stdlib.h>
#include "kernels.h"
#include "tc_ptx.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 32;
constexpr int UMMA_K = 8;
constexpr int NS = 3;                         // pipeline stages
constexpr int TILE_BYTES = BM * BK * 4;       // 16 KB (A and W tiles have the same shape)
constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = 128;

struct __align__(64) TmaMaps { CUtensorMap a, w, wlo; };

using namespace d4tc;
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) { mbar_wait_cta(bar, parity); }
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) { return make_desc_sw128(smem_addr); }
// instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=tf32, K-major both, N>>3 at bit 17, M>>4 at bit 24
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

struct EpiArgs {
    float* C; long long ldc; int M, N;
    const float* bias; const float* row_scale; const float* residual; long long ldr;
    int act; RowMap cmap;
    int a_grp;          // 0: A through a 2D map; else 3D map (K, grp, M/grp)
    int nkb;            // K blocks
};

// ---------------------------------------------------------------- kernel
template <int TERMS>
__global__ void __launch_bounds__(NUM_THREADS) gemm_tc_kernel(const __grid_constant__ TmaMaps maps, const EpiArgs e) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // layout: [stage][A | (A_lo) | W | (W_lo)] tiles, then barriers
    constexpr int TILES_PER_STAGE = (TERMS == 3) ? 4 : 2;
    constexpr int STAGE_BYTES = TILES_PER_STAGE * TILE_BYTES;
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NS * STAGE_BYTES);
    // bars: [0,NS) full, [NS,2NS) empty, [2NS,3NS) split_done, [3NS] tmem_full ; then the TMEM base address word
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * NS + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
    const int nkb = e.nkb;

    auto bar = [&](int i) { return smem_u32(&bars[i]); };
    auto tile = [&](int stage, int which) { return smem + stage * STAGE_BYTES + which * TILE_BYTES; };
    constexpr int T_A = 0, T_ALO = 1, T_W = (TERMS == 3) ? 2 : 1, T_WLO = 3;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.w) : "memory");
        if (TERMS == 3) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.wlo) : "memory");
        for (int s = 0; s < NS; ++s) { mbar_init(bar(s), 1); mbar_init(bar(NS + s), 1); mbar_init(bar(2 * NS + s), 128); }
        mbar_init(bar(3 * NS), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    if (warp == 0) {
        // ================= TMA producer
        if (elect_one()) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % NS; const uint32_t ph = (kb / NS) & 1;
                mbar_wait(bar(NS + s), ph ^ 1);
                mbar_expect_tx(bar(s), (TERMS == 3 ? 3 : 2) * TILE_BYTES);
                if (e.a_grp == 0) tma_load_2d(smem_u32(tile(s, T_A)), &maps.a, bar(s), kb * BK, m0);
                else              tma_load_3d(smem_u32(tile(s, T_A)), &maps.a, bar(s), kb * BK, 0, m0 / e.a_grp);
                tma_load_2d(smem_u32(tile(s, T_W)), &maps.w, bar(s), kb * BK, n0);
                if (TERMS == 3) tma_load_2d(smem_u32(tile(s, T_WLO)), &maps.wlo, bar(s), kb * BK, n0);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer
        if (elect_one()) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % NS; const uint32_t ph = (kb / NS) & 1;
                mbar_wait(bar((TERMS == 3 ? 2 * NS : 0) + s), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t da = make_desc(smem_u32(tile(s, T_A))), dw = make_desc(smem_u32(tile(s, T_W)));
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint64_t koff = (uint64_t)((k * UMMA_K * 4) >> 4);
                    const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
                    if (TERMS == 3) {
                        const uint64_t dalo = make_desc(smem_u32(tile(s, T_ALO))), dwlo = make_desc(smem_u32(tile(s, T_WLO)));
                        umma_tf32(tmem_base, dalo + koff, dw + koff, IDESC, acc);
                        umma_tf32(tmem_base, da + koff, dwlo + koff, IDESC, 1u);
                        umma_tf32(tmem_base, da + koff, dw + koff, IDESC, 1u);
                    } else {
                        umma_tf32(tmem_base, da + koff, dw + koff, IDESC, acc);
                    }
                }
                umma_commit(bar(NS + s));              // frees the stage when these MMAs retire
            }
            umma_commit(bar(3 * NS));                  // accumulator complete
        }
    } else {
        // ================= epilogue warps (and, for tf32x3, the in-smem A splitter)
        const int et = threadIdx.x - 64;               // 0..127
        if (TERMS == 3) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % NS; const uint32_t ph = (kb / NS) & 1;
                mbar_wait(bar(s), ph);
                float4* a = reinterpret_cast<float4*>(tile(s, T_A));
                float4* alo = reinterpret_cast<float4*>(tile(s, T_ALO));
#pragma unroll
                for (int j = 0; j < TILE_BYTES / 16 / 128; ++j) {
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
                mbar_arrive(bar(2 * NS + s));
            }
        }
        mbar_wait(bar(3 * NS), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int quarter = warp & 3;                  // TMEM lane quarter this warp may access
        const int m = m0 + quarter * 32 + lane;
        const bool row_ok = m < e.M;
        const long long crow = row_ok ? e.cmap(m) : 0;
        const float rs = (row_ok && e.row_scale) ? e.row_scale[m] : 1.f;
        const bool glu = (e.act == D4_ACT_GLU_SILU || e.act == D4_ACT_GLU_GELU);
        float* crow_ptr = e.C + crow * e.ldc;
        const float* rrow_ptr = e.residual ? e.residual + crow * e.ldr : nullptr;
        const bool vec_ok = ((e.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(e.C) & 15) == 0) &&
                            (!e.residual || (((e.ldr & 3) == 0) && ((reinterpret_cast<uintptr_t>(e.residual) & 15) == 0)));
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
            if (!row_ok) continue;
            const int nb = n0 + c0;
            if (nb >= e.N) continue;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float t = v[j] * rs;
                if (e.bias && nb + j < e.N) t += e.bias[nb + j];
                v[j] = t;
            }
            if (glu) {
                float o[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float g = v[2 * j + 1];
                    o[j] = v[2 * j] * ((e.act == D4_ACT_GLU_SILU) ? siluf_(g) : geluf_(g));
                }
                const int ob = nb >> 1, on = e.N >> 1;
                if (vec_ok && ob + 16 <= on) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(crow_ptr + ob + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) if (ob + j < on) crow_ptr[ob + j] = o[j];
                }
            } else {
                if (vec_ok && nb + 32 <= e.N) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        if (rrow_ptr) { const float4 r = *reinterpret_cast<const float4*>(rrow_ptr + nb + j); o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w; }
                        *reinterpret_cast<float4*>(crow_ptr + nb + j) = o;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (nb + j < e.N) { float t = v[j]; if (rrow_ptr) t += rrow_ptr[nb + j]; crow_ptr[nb + j] = t; }
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------- host side
// rows x K fp32 matrix, row stride ld floats, box 128 rows x 32 floats, 128B swizzle
int encode_2d(CUtensorMap* map, const float* base, long long rows, long long K, long long ld) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return d4_fail("cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {BK, BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return d4_fail("cuTensorMapEncodeTiled(2d rows=%lld K=%lld ld=%lld) failed: %d", rows, K, ld, (int)r);
    return 0;
}
// row-mapped A: compact row m -> (m / grp) * gstride + goff + (m % grp): 3D map (K, grp, M/grp), box (32, grp, 128/grp)
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

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int TERMS>
int launch(const GemmArgs& g, cudaStream_t stream) {
    TmaMaps maps; memset(&maps, 0, sizeof(maps));
    if (g.amap.grp == 0) { int rc = encode_2d(&maps.a, g.A, g.M, g.K, g.lda); if (rc) return rc; }
    else { int rc = encode_3d(&maps.a, g.A, g.M, g.K, g.lda, g.amap); if (rc) return rc; }
    { int rc = encode_2d(&maps.w, g.W, g.N, g.K, g.ldw); if (rc) return rc; }
    if (TERMS == 3) { int rc = encode_2d(&maps.wlo, g.W_lo, g.N, g.K, g.ldw); if (rc) return rc; }
    EpiArgs e;
    e.C = g.C; e.ldc = g.ldc; e.M = g.M; e.N = g.N; e.bias = g.bias; e.row_scale = g.row_scale; e.residual = g.residual; e.ldr = g.ldr;
    e.act = g.act; e.cmap = g.cmap; e.a_grp = g.amap.grp; e.nkb = (g.K + BK - 1) / BK;
    constexpr int smem = NS * ((TERMS == 3) ? 4 : 2) * TILE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
    static bool configured = false;
    if (!configured) {
        D4_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<TERMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM);
    gemm_tc_kernel<TERMS><<<grid, NUM_THREADS, smem, stream>>>(maps, e);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace

int d4_gemm_tc_supported(const GemmArgs& g) {
    if (g.transA || g.transW) return 0;
    if (g.M < 1 || g.N < 1 || g.K < BK || (g.K & 3)) return 0;
    if ((g.lda & 3) || (g.ldw & 3) || !al16(g.A) || !al16(g.W)) return 0;
    if ((g.M + BM - 1) / BM > 65535) return 0;
    if (g.amap.grp != 0) {
        if (BM % g.amap.grp != 0 || g.M % g.amap.grp != 0 || g.amap.grp > 256) return 0;
        if (((long long)g.amap.goff * g.lda) & 3) return 0;
    }
    if ((g.act == D4_ACT_GLU_SILU || g.act == D4_ACT_GLU_GELU) && (g.N & 1)) return 0;
    if (g.act == D4_ACT_SILU) return 0;
    return 1;
}

// Default: the CTA-pair kernel in gemm_tc3.cu (single-CTA persistent kernel of gemm_tc2.cu when M fits one CTA's rows).
// D4_GEMM_V=2 forces gemm_tc2.cu, D4_GEMM_V=1 this one-tile-per-CTA kernel; D4_GEMM_BN=128|256 pins the N tile.
// 1 if a GEMM with more than 128 rows goes to the CTA-pair kernel (the engine may then fuse the RMS statistics into it)
int d4_gemm_pair_default(void) {
    const char* v = getenv("D4_GEMM_V");
    return (!v || atoi(v) == 3) ? 1 : 0;
}

int d4_gemm_tc(const GemmArgs& g, int terms, cudaStream_t stream) {
    if (!d4_gemm_tc_supported(g)) return d4_fail("gemm_tc: unsupported shape / alignment");
    if (terms == 3 && (!g.W_lo || !al16(g.W_lo))) return d4_fail("gemm_tc: tf32x3 needs a 16-byte aligned W_lo");
    static int version = -1, bn = 0;
    if (version < 0) {
        const char* v = getenv("D4_GEMM_V"); version = v ? atoi(v) : 3;
        const char* b = getenv("D4_GEMM_BN"); bn = b ? atoi(b) : 0;
    }
    if (version == 3 && g.M > BM) return d4_gemm_tc3(g, terms == 3 ? 3 : 1, bn, stream);
    if (g.rs_mode || g.ss_out) return d4_fail("gemm_tc: sum-of-squares row statistics need the CTA-pair kernel (M > 128, D4_GEMM_V unset)");
    if (version != 1) return d4_gemm_tc2(g, terms == 3 ? 3 : 1, bn, stream);
    if (terms == 3) return launch<3>(g, stream);
    return launch<1>(g, stream);
}
