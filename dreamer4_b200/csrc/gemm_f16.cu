// The dominant kernel of the `f16x3` engine mode (the bench default since round 2): the CTA-pair tcgen05 GEMM of gemm_tc3.cu with
// the three split terms on kind::f16 instead of kind::tf32.
//
// Why: fp16 carries the same 11-bit significand as TF32 and kind::f16 issues at twice the TF32 rate, so
//     D = a_hi.w_hi + a_lo.w_hi + a_hi.w_lo,   hi = fp16(x), lo = fp16(x - hi)
// costs 1.5 TF32-MMA units per product instead of 3 at the accuracy of the 3xTF32 split (scripts/split_precision_study.py,
// profiles/r1_split_precision_study.txt: 2.5e-8 .. 1.3e-7 of sum|a.w| against 3.4e-8 .. 1.9e-7), PROVIDED the operands sit
// inside fp16's exponent range.  That is arranged with exact power-of-two scales:
//   * each row of A is multiplied by p[m] = 2^round(log2(rstd[m])) in the splitter when the caller runs in rs_mode 1 (the
//     epilogue applies rstd anyway, so it knows the row's rms), 1 otherwise; the epilogue multiplies by rs / p;
//   * the weights arrive pre-scaled to rms ~ 1 (W' = q W, q a power of two chosen once per matrix on the host) as two fp16
//     arrays (hi, lo); the epilogue multiplies by w_scale = 1 / q.
//
// Differences from gemm_tc3_kernel<3, BN, 32>:
//   K step 64 elements (one 128-byte swizzled fp16 row) - the same 12 MMAs per stage now cover twice the K;
//   A still arrives as fp32 through TMA (two 32-column boxes per stage); the eight splitter warps read the whole 32 KB into
//   registers, meet at a named barrier, and write the fp16 hi tile and lo tile back IN PLACE (16 KB each) in the 128B-swizzled
//   K-major layout; W hi / lo are fp16 in global memory and land by TMA directly;
//   instruction descriptor a_format = b_format = F16 (0), UMMA_K = 16;
//   the residual chunk loads of the epilogue are issued early (see the epilogue).
// Roles, barriers and TMEM double buffering are those of gemm_tc3.cu.  Measured: DESIGN.md section 5 (K5), profiles/r2_gemm_*.txt.
#include <cuda.h>
#include <cuda_fp16.h>
#include <string.h>
#include <stdlib.h>
#include <algorithm>
#include "kernels.h"
#include "tc_ptx.cuh"

namespace {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int NUM_THREADS = 512;          // warp 0 TMA, 1 MMA, (2-3 idle), 4-7 epilogue, 8-15 operand splitter
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;

struct __align__(64) TmaMapsH { CUtensorMap a, w, wlo, c, r; };

using d4tc::smem_u32; using d4tc::elect_one; using d4tc::mbar_init; using d4tc::mbar_expect_tx; using d4tc::tma_load_2d; using d4tc::tma_load_3d;
using d4tc::tmem_ld32; using d4tc::EncodeTiledFn; using d4tc::get_encode; using d4tc::make_desc_sw128;
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (int spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1 << 27)) __trap();          // never hang the box on a protocol bug
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar & PEER_MASK), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ float silu_fast(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_c), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
// p = 2^round(log2(x)) for a positive normal x (rounds the exponent by adding half the mantissa range), and its exact inverse
__device__ __forceinline__ float pow2_near(float x) { return __uint_as_float((__float_as_uint(x) + 0x00400000u) & 0x7F800000u); }
__device__ __forceinline__ float pow2_inv(float p) { return __uint_as_float(0x7F000000u - __float_as_uint(p)); }

struct EpiArgsH {
    float* C; long long ldc; int M, N;
    const float* bias; const float* row_scale; const float* residual; long long ldr;
    int act; RowMap cmap;
    int a_grp, nkb, n_tiles_m, n_tiles_n;
    int rs_mode, kdim;  // rs_mode 1: row_scale holds the sum of squares over kdim columns -> rsqrt(ss / kdim + eps); rows pre-scaled by 2^k
    float* ss_out;
    float w_scale;      // 1 / q of the pre-scaled weights
    int dbg;            // ablation switches for scripts/gemm_bench.py (results are garbage when set): 1 no epilogue, 2 no split, 4 no MMA, 8 / 16 same A / W tile; 64 residual chunk loads issued late (results stay exact), 128 no output stores
};

template <int BN> struct CfgH {
    static constexpr int A_BOX = BM * 32 * 4;                           // one fp32 TMA box (32 columns); also one fp16 tile (64 columns)
    static constexpr int A_STAGE = 2 * A_BOX;                           // fp32 [k 0..31 | k 32..63] in, fp16 [hi | lo] after the split
    static constexpr int BNH = BN / 2;
    static constexpr int W_TILE = BNH * BK * 2;                         // fp16
    static constexpr int STAGE = A_STAGE + 2 * W_TILE;                  // 64 KB (BN 256) / 48 KB (BN 128)
    static constexpr int NS = (192 * 1024) / STAGE;                     // 3 / 4
    static constexpr int STG_BYTES = 4 * 2 * 4096;
    static constexpr int NBARS = 4 * NS + 4 + 8;                        // full | fullA | empty | split | tfull[2] | tempty[2] | resid[4][2]
    static constexpr int SMEM = NS * STAGE + STG_BYTES + NBARS * 8 + 64 + 1024;
    // instruction descriptor: D = f32 (1 at bit 4), A = B = f16 (0 at bits 7, 10), K-major, N >> 3 at bit 17, M >> 4 at bit 24 (M = 256)
    static constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
};

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_f16x3_kernel(const __grid_constant__ TmaMapsH maps, const EpiArgsH e) {
    using K = CfgH<BN>;
    constexpr int NS = K::NS;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* stg_all = reinterpret_cast<float*>(smem + NS * K::STAGE);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NS * K::STAGE + K::STG_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + K::NBARS);
    auto bar = [&](int i) { return smem_u32(&bars[i]); };
    constexpr int B_FULL = 0, B_FULLA = NS, B_EMPTY = 2 * NS, B_SPLIT = 3 * NS, B_TFULL = 4 * NS, B_TEMPTY = 4 * NS + 2, B_RES = 4 * NS + 4;
    constexpr int T_A = 0, T_ALO = K::A_BOX, T_W = K::A_STAGE, T_WLO = T_W + K::W_TILE;
    auto tile = [&](int stage, int off) { return smem + stage * K::STAGE + off; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int nkb = e.nkb;
    const int total_tiles = e.n_tiles_m * e.n_tiles_n;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.wlo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.c) : "memory");
        if (e.residual) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.r) : "memory");
        for (int s = 0; s < NS; ++s) {
            mbar_init(bar(B_FULL + s), 1); mbar_init(bar(B_FULLA + s), 1); mbar_init(bar(B_EMPTY + s), 1); mbar_init(bar(B_SPLIT + s), 2);
        }
        for (int b = 0; b < 2; ++b) { mbar_init(bar(B_TFULL + b), 1); mbar_init(bar(B_TEMPTY + b), 8); }
        for (int b = 0; b < 8; ++b) mbar_init(bar(B_RES + b), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    if (warp == 0) {
        // ================= TMA producer (both CTAs): A as two fp32 boxes on this CTA's barrier, W hi / lo (fp16) on the leader's
        if (elect_one()) {
            uint32_t kc = 0;
            for (int t = cluster_id; t < total_tiles; t += n_clusters) {
                // dbg 8 / 16 (timing experiments): every cluster fetches the A / the W tile of tile 0 - how much of the data-movement
                // floor is L2 slice bandwidth (identical lines requested by all SMs) and how much is delivery to the SMs
                const int m0 = ((e.dbg & 8) ? 0 : (t / e.n_tiles_n) * (2 * BM)) + (int)rank * BM;
                const int n0 = ((e.dbg & 16) ? 0 : (t % e.n_tiles_n) * BN) + (int)rank * K::BNH;
                for (int kb = 0; kb < nkb; ++kb, ++kc) {
                    const int s = kc % NS; const uint32_t ph = (kc / NS) & 1;
                    mbar_wait(bar(B_EMPTY + s), ph ^ 1);
                    mbar_expect_tx(bar(B_FULLA + s), K::A_STAGE);
                    for (int h = 0; h < 2; ++h) {
                        if (e.a_grp == 0) tma_load_2d(smem_u32(tile(s, T_A + h * K::A_BOX)), &maps.a, bar(B_FULLA + s), kb * BK + h * 32, m0);
                        else              tma_load_3d(smem_u32(tile(s, T_A + h * K::A_BOX)), &maps.a, bar(B_FULLA + s), kb * BK + h * 32, 0, m0 / e.a_grp);
                    }
                    if (rank == 0) mbar_expect_tx(bar(B_FULL + s), 4 * K::W_TILE);
                    tma_load_2d_pair(smem_u32(tile(s, T_W)), &maps.w, bar(B_FULL + s), kb * BK, n0);
                    tma_load_2d_pair(smem_u32(tile(s, T_WLO)), &maps.wlo, bar(B_FULL + s), kb * BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA only)
        if (rank == 0 && elect_one()) {
            uint32_t kc = 0, ac = 0;
            for (int t = cluster_id; t < total_tiles; t += n_clusters, ++ac) {
                const int buf = ac & 1; const uint32_t aph = (ac >> 1) & 1;
                mbar_wait(bar(B_TEMPTY + buf), aph ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_c = tmem_base + (uint32_t)(buf * BN);
                for (int kb = 0; kb < nkb; ++kb, ++kc) {
                    const int s = kc % NS; const uint32_t ph = (kc / NS) & 1;
                    mbar_wait(bar(B_FULL + s), ph);
                    mbar_wait(bar(B_SPLIT + s), ph);                     // both CTAs' fp16 hi / lo tiles of A are in place
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t da = make_desc_sw128(smem_u32(tile(s, T_A))), dalo = make_desc_sw128(smem_u32(tile(s, T_ALO)));
                    const uint64_t dw = make_desc_sw128(smem_u32(tile(s, T_W))), dwlo = make_desc_sw128(smem_u32(tile(s, T_WLO)));
                    if (!(e.dbg & 4))
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {             // small terms first, the hi*hi product last
                        const uint64_t koff = (uint64_t)((k * UMMA_K * 2) >> 4);
                        umma_f16_pair(tmem_c, dalo + koff, dw + koff, K::IDESC, (kb > 0 || k > 0) ? 1u : 0u);
                        umma_f16_pair(tmem_c, da + koff, dwlo + koff, K::IDESC, 1u);
                        umma_f16_pair(tmem_c, da + koff, dw + koff, K::IDESC, 1u);
                    }
                    umma_commit_pair(bar(B_EMPTY + s));
                }
                umma_commit_pair(bar(B_TFULL + buf));
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ================= epilogue of this CTA's 128 rows (gemm_tc3.cu's, with the row scale corrected for the two pre-scales)
        const int quarter = warp & 3, ew = warp - 4;
        unsigned char* ebuf = reinterpret_cast<unsigned char*>(stg_all) + ew * 2 * 4096;
        const uint32_t ebuf_u32 = smem_u32(ebuf);
        const bool glu = (e.act == D4_ACT_GLU_SILU || e.act == D4_ACT_GLU_GELU);
        const bool has_res = (e.residual != nullptr);
        const float* __restrict__ biasp = e.bias;
        const uint32_t tempty_leader = bar(B_TEMPTY) & PEER_MASK;
        const int grp = e.cmap.grp;
        const int n_out = glu ? (e.N >> 1) : e.N;
        const int in_per_chunk = glu ? 64 : 32;
        const uint32_t swz = (uint32_t)(lane & 7);
        uint32_t g = 0, rph0 = 0, rph1 = 0;
        auto res_load = [&](int buf, int col0, int row0) {     // lane 0 only
            const uint32_t rb = bar(B_RES + ew * 2 + buf);
            mbar_expect_tx(rb, 4096);
            if (grp == 0) tma_load_2d(ebuf_u32 + buf * 4096, &maps.r, rb, col0, row0);
            else          tma_load_3d(ebuf_u32 + buf * 4096, &maps.r, rb, col0, 0, row0 / grp);
        };
        // The residual tile is added in place in the chunk buffers (TMA load -> += -> TMA store).  With two 4 KB buffers per warp a
        // chunk's load can only be issued once the store two chunks back has left its buffer; issued after the current chunk's
        // store (round 1) its whole latency was exposed on every chunk, which made the epilogue of the N = 512, K = 256 / 512
        // products (attention out, pool out) longer than their mainloop.  Now the next chunk's load is issued a quarter into the
        // current chunk's arithmetic - the previous store has been read by then - also across the tile boundary: attention out
        // 89 -> 79 us, pool out 87 -> 71 us at 30720 rows (profiles/r2_gemm_residual_epilogue.txt; dbg 64 restores the late issue).
        // (Tried on top and dropped: prefetching the next tile's residual rows into L2 with cp.async.bulk.prefetch.tensor - 6-8 %
        // slower; an epilogue without staging, every lane storing its row's 32 columns as eight 16-byte st.global and reading the
        // residual the same way - its arithmetic is cheaper (no-store timing 141 vs 159 us on the qkv shape) but 32 half-sector
        // stores per instruction cost far more than the TMA store: 199 vs 164 us, 30.1 k vs 32.3 k frames/s in situ.)
        const bool res_early = has_res && !(e.dbg & 64);
        bool pre_loaded = false;                                // lane 0: the first chunk of the coming tile is already in flight
        auto tile_rows = [&](int tt, int& rb, int& o0) {
            rb = (tt / e.n_tiles_n) * (2 * BM) + (int)rank * BM + quarter * 32;
            const int nn = (tt % e.n_tiles_n) * BN;
            o0 = glu ? (nn >> 1) : nn;
        };
        uint32_t ac = 0;
        for (int t = cluster_id; t < total_tiles; t += n_clusters, ++ac) {
            const int m0 = (t / e.n_tiles_n) * (2 * BM) + (int)rank * BM, n0 = (t % e.n_tiles_n) * BN;
            const int buf_acc = ac & 1; const uint32_t aph = (ac >> 1) & 1;
            const int rbase = m0 + quarter * 32;
            const int mrow = rbase + lane;
            float rs = (mrow < e.M && e.row_scale) ? e.row_scale[mrow] : 1.f;
            if (e.rs_mode && e.row_scale) { rs = rsqrtf(rs / (float)e.kdim + D4_RMS_EPS); rs *= pow2_inv(pow2_near(rs)); }   // rows were fed as a * 2^k
            rs *= e.w_scale;
            float ss_part = 0.f;
            const uint32_t tmem_c = tmem_base + (uint32_t)(buf_acc * BN) + ((uint32_t)(quarter * 32) << 16);
            const int out0 = glu ? (n0 >> 1) : n0;
            const bool rows_ok = rbase < e.M;
            if (has_res && rows_ok && out0 < n_out && lane == 0 && !pre_loaded) { bulk_wait_read_1(); res_load(g & 1, out0, rbase); }
            pre_loaded = false;
            mbar_wait(bar(B_TFULL + buf_acc), aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (rows_ok && !(e.dbg & 1)) {
                auto bias_regs = [&](int nb, float& lo, float& hi) {
                    lo = (biasp && nb + lane < e.N) ? __ldg(biasp + nb + lane) : 0.f;
                    hi = (biasp && glu && nb + 32 + lane < e.N) ? __ldg(biasp + nb + 32 + lane) : 0.f;
                };
                float b_lo, b_hi, bn_lo = 0.f, bn_hi = 0.f;
                bias_regs(n0, b_lo, b_hi);
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += in_per_chunk) {
                    const int oc = out0 + (glu ? (c0 >> 1) : c0);
                    if (oc >= n_out) break;
                    const int buf = g & 1;
                    const int nb = n0 + c0;
                    if (c0 + in_per_chunk < BN) bias_regs(nb + in_per_chunk, bn_lo, bn_hi);
                    float v[32], w[32];
                    tmem_ld32(tmem_c + (uint32_t)c0, v);
                    if (glu) tmem_ld32(tmem_c + (uint32_t)(c0 + 32), w);
                    if (lane == 0) bulk_wait_read_1();
                    __syncwarp();
                    if (has_res) {
                        if (buf == 0) { mbar_wait(bar(B_RES + ew * 2), rph0); rph0 ^= 1; }
                        else          { mbar_wait(bar(B_RES + ew * 2 + 1), rph1); rph1 ^= 1; }
                    }
                    unsigned char* rowp = ebuf + buf * 4096 + lane * 128;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        if (q == 2 && res_early && lane == 0) {          // next chunk's residual: its buffer's store (the previous chunk) has been read by now
                            const int oc_next = oc + 32;
                            if (c0 + in_per_chunk < BN && oc_next < n_out) { bulk_wait_read_0(); res_load(buf ^ 1, oc_next, rbase); }
                            else if (t + n_clusters < total_tiles) {
                                int rb, o0; tile_rows(t + n_clusters, rb, o0);
                                if (rb < e.M && o0 < n_out) { bulk_wait_read_0(); res_load(buf ^ 1, o0, rb); pre_loaded = true; }
                            }
                        }
                        float4 o;
                        if (!glu) {
                            const float bx = __shfl_sync(0xffffffffu, b_lo, 4 * q), by = __shfl_sync(0xffffffffu, b_lo, 4 * q + 1);
                            const float bz = __shfl_sync(0xffffffffu, b_lo, 4 * q + 2), bw = __shfl_sync(0xffffffffu, b_lo, 4 * q + 3);
                            o.x = fmaf(v[4 * q], rs, bx); o.y = fmaf(v[4 * q + 1], rs, by);
                            o.z = fmaf(v[4 * q + 2], rs, bz); o.w = fmaf(v[4 * q + 3], rs, bw);
                        } else {
                            const float* src = (q < 4) ? (v + 8 * q) : (w + 8 * (q - 4));
                            const float bsrc = (q < 4) ? b_lo : b_hi;
                            float r4[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float x = fmaf(src[2 * k], rs, __shfl_sync(0xffffffffu, bsrc, (8 * q + 2 * k) & 31));
                                const float gt = fmaf(src[2 * k + 1], rs, __shfl_sync(0xffffffffu, bsrc, (8 * q + 2 * k + 1) & 31));
                                r4[k] = x * ((e.act == D4_ACT_GLU_SILU) ? silu_fast(gt) : geluf_(gt));
                            }
                            o = make_float4(r4[0], r4[1], r4[2], r4[3]);
                        }
                        float4* dst = reinterpret_cast<float4*>(rowp + (((uint32_t)q ^ swz) << 4));
                        if (has_res) { const float4 r = *dst; o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w; }
                        *dst = o;
                        ss_part = fmaf(o.x, o.x, fmaf(o.y, o.y, fmaf(o.z, o.z, fmaf(o.w, o.w, ss_part))));
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        if (e.dbg & 128) {}          // timing experiment: the epilogue without its stores
                        else if (grp == 0) tma_store_2d(&maps.c, ebuf_u32 + buf * 4096, oc, rbase);
                        else          tma_store_3d(&maps.c, ebuf_u32 + buf * 4096, oc, 0, rbase / grp);
                        bulk_commit();
                        const int oc_next = oc + 32;
                        if (has_res && !res_early && c0 + in_per_chunk < BN && oc_next < n_out) { bulk_wait_read_1(); res_load(buf ^ 1, oc_next, rbase); }
                    }
                    b_lo = bn_lo; b_hi = bn_hi;
                    ++g;
                }
            }
            if (e.ss_out && mrow < e.M) atomicAdd(e.ss_out + mrow, ss_part);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_leader + (uint32_t)buf_acc * 8u);
        }
        if (lane == 0) bulk_wait_all();
    } else if (warp >= 8) {
        // ================= A splitter: this CTA's fp32 A stage -> fp16 hi tile + fp16 lo tile, in place
        // unit u = (row r, 16-byte fp16 chunk c): the row's elements 8c..8c+7 = fp32 chunks 2(c&3), 2(c&3)+1 of box c>>2; both
        // layouts are 128-byte rows with chunk j stored at j ^ (r & 7).  256 threads x 4 units = 128 rows x 8 chunks.
        const int et = threadIdx.x - 256;          // 0..255
        const uint32_t split_leader = bar(B_SPLIT) & PEER_MASK;
        uint32_t kc = 0;
        for (int t = cluster_id; t < total_tiles; t += n_clusters) {
            const int m0 = (t / e.n_tiles_n) * (2 * BM) + (int)rank * BM;
            float pre[4];                           // power-of-two row pre-scales of this thread's four rows
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int mrow = m0 + ((et + 256 * j) >> 3);
                float p = 1.f;
                if (e.rs_mode && e.row_scale && mrow < e.M) p = pow2_near(rsqrtf(e.row_scale[mrow] / (float)e.kdim + D4_RMS_EPS));
                pre[j] = p;
            }
            for (int kb = 0; kb < nkb; ++kb, ++kc) {
                const int s = kc % NS; const uint32_t ph = (kc / NS) & 1;
                mbar_wait(bar(B_FULLA + s), ph);
                if (e.dbg & 2) { if (et == 0) mbar_arrive_cluster(split_leader + (uint32_t)s * 8u); continue; }
                unsigned char* base = tile(s, T_A);
                float4 x[4][2];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int u = et + 256 * j, r = u >> 3, c = u & 7;
                    const unsigned char* rowp = base + (c >> 2) * K::A_BOX + r * 128;
                    const uint32_t q0 = 2u * (uint32_t)(c & 3), sw = (uint32_t)(r & 7);
                    x[j][0] = *reinterpret_cast<const float4*>(rowp + ((q0 ^ sw) << 4));
                    x[j][1] = *reinterpret_cast<const float4*>(rowp + (((q0 + 1u) ^ sw) << 4));
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");          // every fp32 word has been read: the stage may be overwritten
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int u = et + 256 * j, r = u >> 3, c = u & 7;
                    const float p = pre[j];
                    const float f[8] = {x[j][0].x * p, x[j][0].y * p, x[j][0].z * p, x[j][0].w * p, x[j][1].x * p, x[j][1].y * p, x[j][1].z * p, x[j][1].w * p};
                    __half2 hi[4], lo[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const __half h0 = __float2half_rn(f[2 * k]), h1 = __float2half_rn(f[2 * k + 1]);
                        hi[k] = __halves2half2(h0, h1);
                        lo[k] = __halves2half2(__float2half_rn(f[2 * k] - __half2float(h0)), __float2half_rn(f[2 * k + 1] - __half2float(h1)));
                    }
                    const uint32_t off = (uint32_t)r * 128u + ((((uint32_t)c) ^ (uint32_t)(r & 7)) << 4);
                    *reinterpret_cast<uint4*>(base + off) = *reinterpret_cast<const uint4*>(hi);
                    *reinterpret_cast<uint4*>(base + K::A_BOX + off) = *reinterpret_cast<const uint4*>(lo);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (et == 0) mbar_arrive_cluster(split_leader + (uint32_t)s * 8u);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
    }
}

// ---------------------------------------------------------------- host side
int encode_a32(CUtensorMap* map, const float* base, long long M, long long K, long long ld, const RowMap& rm) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return d4_fail("cuTensorMapEncodeTiled is not available from the driver");
    CUresult r;
    if (rm.grp == 0) {
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
        cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
        cuuint32_t box[2] = {32, (cuuint32_t)BM};
        cuuint32_t estr[2] = {1, 1};
        r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rm.grp, (cuuint64_t)(M / rm.grp)};
        cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)rm.gstride * ld * 4};
        cuuint32_t box[3] = {32, (cuuint32_t)rm.grp, (cuuint32_t)(BM / rm.grp)};
        cuuint32_t estr[3] = {1, 1, 1};
        r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base + (long long)rm.goff * ld), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) return d4_fail("gemm_f16: cuTensorMapEncodeTiled(A M=%lld K=%lld ld=%lld) failed: %d", M, K, ld, (int)r);
    return 0;
}
// fp16 weights (N, K) with leading dimension ld (in fp16 elements, a multiple of 8): 64-column boxes = 128-byte swizzled rows
int encode_w16(CUtensorMap* map, const void* base, long long N, long long K, long long ld, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return d4_fail("cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return d4_fail("gemm_f16: cuTensorMapEncodeTiled(W N=%lld K=%lld ld=%lld) failed: %d", N, K, ld, (int)r);
    return 0;
}
int encode_out32(CUtensorMap* map, const float* base, long long M, long long N, long long ld, const RowMap& rm) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return d4_fail("cuTensorMapEncodeTiled is not available from the driver");
    CUresult r;
    if (rm.grp == 0) {
        cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
        cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
        cuuint32_t box[2] = {32, 32};
        cuuint32_t estr[2] = {1, 1};
        r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)rm.grp, (cuuint64_t)(M / rm.grp)};
        cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)rm.gstride * ld * 4};
        cuuint32_t box[3] = {32, (cuuint32_t)rm.grp, (cuuint32_t)(32 / rm.grp)};
        cuuint32_t estr[3] = {1, 1, 1};
        r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base + (long long)rm.goff * ld), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) return d4_fail("gemm_f16: cuTensorMapEncodeTiled(out M=%lld N=%lld ld=%lld) failed: %d", M, N, ld, (int)r);
    return 0;
}
inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// What the kernel can take: the TMA epilogue only (16-byte aligned rows, row maps whose groups tile the 32-row warp slices),
// fp16 weight rows of ldw elements (a multiple of 8 = 16 bytes), sum-of-squares output on identity rows without GLU.
bool operands_ok(const GemmArgs& g, const void* whi, const void* wlo) {
    const bool glu = (g.act == D4_ACT_GLU_SILU || g.act == D4_ACT_GLU_GELU);
    const int grp = g.cmap.grp;
    if (g.M < 1 || g.N < 1 || g.K < 1 || (glu && (g.N & 1)) || g.act == D4_ACT_SILU) return false;
    bool ok = whi && wlo && al16(g.A) && (g.lda % 4 == 0) && al16(whi) && al16(wlo) && (g.ldw % 8 == 0) && al16(g.C) && (g.ldc % 4 == 0) && (!g.bias || al16(g.bias)) &&
              (grp == 0 || (32 % grp == 0 && g.M % grp == 0 && (((long long)g.cmap.goff * g.ldc) % 4 == 0) && (((long long)g.cmap.gstride * g.ldc) % 4 == 0))) &&
              (g.amap.grp == 0 || (BM % g.amap.grp == 0 && g.M % g.amap.grp == 0 && g.amap.grp <= 256 && (((long long)g.amap.goff * g.lda) % 4 == 0))) &&
              !g.transA && !g.transW;
    if (g.residual) ok = ok && al16(g.residual) && (g.ldr % 4 == 0) &&
                         (grp == 0 || ((((long long)g.cmap.goff * g.ldr) % 4 == 0) && (((long long)g.cmap.gstride * g.ldr) % 4 == 0)));
    if (g.ss_out) ok = ok && grp == 0 && !glu && (g.N % 4) == 0;
    return ok;
}

int g_dbg = [] { const char* v = getenv("D4_GEMM_F16_DBG"); return v ? atoi(v) & 255 : 0; }();          // d4_debug_set("gemm_f16", bits): 1 no epilogue, 2 no operand split, 4 no MMA (timing ablations only)

template <int BN>
int launch_h(const GemmArgs& g, float w_scale, cudaStream_t stream) {
    using K = CfgH<BN>;
    static bool configured = false;
    static int maxc = 0;
    if (!configured) {
        D4_CUDA_OK(cudaFuncSetAttribute(gemm_f16x3_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM));
        cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(2, 1, 1); cfg.blockDim = dim3(NUM_THREADS, 1, 1); cfg.dynamicSmemBytes = K::SMEM;
        cudaLaunchAttribute attr; attr.id = cudaLaunchAttributeClusterDimension; attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        if (cudaOccupancyMaxActiveClusters(&maxc, gemm_f16x3_kernel<BN>, &cfg) != cudaSuccess || maxc <= 0) { cudaGetLastError(); maxc = -1; }
        configured = true;
    }
    if (maxc <= 0) return d4_fail("gemm_f16: no CTA pair of %d bytes of shared memory can be scheduled on this device", K::SMEM);
    const bool glu = (g.act == D4_ACT_GLU_SILU || g.act == D4_ACT_GLU_GELU);
    if (!operands_ok(g, g.W, g.W_lo)) return d4_fail("gemm_f16: operand alignment / row map not supported by this kernel");
    TmaMapsH maps; memset(&maps, 0, sizeof(maps));
    { int rc = encode_a32(&maps.a, g.A, g.M, g.K, g.lda, g.amap); if (rc) return rc; }
    { int rc = encode_w16(&maps.w, g.W, g.N, g.K, g.ldw, K::BNH); if (rc) return rc; }
    { int rc = encode_w16(&maps.wlo, g.W_lo, g.N, g.K, g.ldw, K::BNH); if (rc) return rc; }
    { int rc = encode_out32(&maps.c, g.C, g.M, glu ? g.N / 2 : g.N, g.ldc, g.cmap); if (rc) return rc; }
    if (g.residual) { int rc = encode_out32(&maps.r, g.residual, g.M, g.N, g.ldr, g.cmap); if (rc) return rc; }
    EpiArgsH e; memset(&e, 0, sizeof(e));
    e.C = g.C; e.ldc = g.ldc; e.M = g.M; e.N = g.N; e.bias = g.bias; e.row_scale = g.row_scale; e.residual = g.residual; e.ldr = g.ldr;
    e.act = g.act; e.cmap = g.cmap; e.a_grp = g.amap.grp; e.nkb = (g.K + BK - 1) / BK;
    e.n_tiles_m = (g.M + 2 * BM - 1) / (2 * BM); e.n_tiles_n = (g.N + BN - 1) / BN;
    e.rs_mode = g.rs_mode; e.kdim = g.K; e.ss_out = g.ss_out; e.w_scale = w_scale; e.dbg = g_dbg;
    const long long tiles = (long long)e.n_tiles_m * e.n_tiles_n;
    const int clusters = (int)std::min<long long>(tiles, maxc);
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * clusters, 1, 1); cfg.blockDim = dim3(NUM_THREADS, 1, 1); cfg.dynamicSmemBytes = K::SMEM; cfg.stream = stream;
    cudaLaunchAttribute attr; attr.id = cudaLaunchAttributeClusterDimension; attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    D4_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_f16x3_kernel<BN>, maps, e));
    D4_COUNT_LAUNCH();
    return 0;
}

}  // namespace

// 1 if the engine may route this GEMM here given the fp16 hi / lo arrays of its weight: the pair kernel's shapes (more rows than
// one CTA's 128, like gemm_tc3.cu), whole 64-column K steps (K = dim, ff_inner_pad, pool width ... of every BASELINE config), and
// the operand rules above.  g.W still points to the fp32 weight; only its leading dimension is read.
void d4_gemm_f16_debug(int bits) { g_dbg = bits & 255; }
int d4_gemm_f16x3_supported(const GemmArgs& g, const void* whi, const void* wlo) {
    return (g.M > BM && (g.K % 32) == 0 && g.K >= BK && operands_ok(g, whi, wlo)) ? 1 : 0;
}

// g.W / g.W_lo point to fp16 arrays (N, ldw) holding hi = fp16(q W), lo = fp16(q W - hi); w_scale = 1 / q.  bn = 128 | 256 (0: by padding).
int d4_gemm_f16x3(const GemmArgs& g, float w_scale, int bn, cudaStream_t stream) {
    if (!g.W_lo) return d4_fail("gemm_f16: needs the low fp16 words of W");
    if (bn == 0) bn = d4_gemm_pair_bn(g.M, g.N);
    return bn == 256 ? launch_h<256>(g, w_scale, stream) : launch_h<128>(g, w_scale, stream);
}
