// Skinny GEMM: C (M, N) = epi(rs[m] * A (M, K) @ W (N, K)^T + bias) for M <= 32 rows - the linear layers of a pass at tiny batches
// (one env / a few dreams: M = B * 15 token rows, or B rows for the heads) and the B-row projections of every pass.
//
// Why a kernel of its own: with a handful of rows the layer is a weight stream (W: 1 - 11 MB, 0.5 FLOP per byte), but a 128-row
// tensor-core tile gives the persistent tcgen05 kernels only N / 128 CTAs to pull it with - 13 SMs for the fused qkv projection,
// 30 us per launch measured at batch 1 (profiles/r2_env_step_launches_b1.txt), 75 % of an env step.  Here N is cut into NW-row
// slabs, one 128-thread CTA per slab (N / NW CTAs: 128 - 680 for the layer shapes of the pass), the four warps of a CTA split K,
// every lane streams W with 128-bit loads and keeps MT x NW exact-fp32 accumulators; A (<= 32 x K floats) is read through L1.
// Exact fp32 FMA: tiny batches get the oracle's arithmetic (reassociated) in every engine precision.
#include "kernels.h"

namespace {

constexpr int NTHR = 128, NWARP = NTHR / 32;

template <int MT, int NW>
__global__ void __launch_bounds__(NTHR) gemm_skinny_kernel(GemmArgs g) {
    __shared__ float part[NWARP][MT][NW];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * NW;
    const int K4 = g.K >> 2;
    float acc[MT][NW];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int j = 0; j < NW; ++j) acc[m][j] = 0.f;
    const float4* wrow[NW];
#pragma unroll
    for (int j = 0; j < NW; ++j) wrow[j] = reinterpret_cast<const float4*>(g.W + (long long)min(n0 + j, g.N - 1) * g.ldw);
    const float4* arow[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) arow[m] = reinterpret_cast<const float4*>(g.A + g.amap(min(m, g.M - 1)) * g.lda);

    for (int q = lane + 32 * warp; q < K4; q += NTHR) {
        float4 w[NW];
#pragma unroll
        for (int j = 0; j < NW; ++j) w[j] = __ldg(wrow[j] + q);
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            const float4 a = __ldg(arow[m] + q);
#pragma unroll
            for (int j = 0; j < NW; ++j)
                acc[m][j] = fmaf(a.x, w[j].x, fmaf(a.y, w[j].y, fmaf(a.z, w[j].z, fmaf(a.w, w[j].w, acc[m][j]))));
        }
    }
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int j = 0; j < NW; ++j) {
            float v = acc[m][j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(D4_FULL, v, o);
            if (lane == 0) part[warp][m][j] = v;
        }
    __syncthreads();
    const bool glu = (g.act == D4_ACT_GLU_SILU || g.act == D4_ACT_GLU_GELU);
    for (int e = threadIdx.x; e < MT * NW; e += NTHR) {
        const int m = e / NW, j = e % NW, n = n0 + j;
        if (m >= g.M || n >= g.N) continue;
        float rs = g.row_scale ? g.row_scale[m] : 1.f;
        if (g.rs_mode && g.row_scale) rs = rsqrtf(rs / (float)g.K + D4_RMS_EPS);
        auto val = [&](int jj) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < NWARP; ++w) s += part[w][m][jj];
            s *= rs;
            if (g.bias) s += g.bias[n0 + jj];
            return s;
        };
        const long long crow = g.cmap(m);
        if (glu) {
            if (j & 1) continue;                       // rows 2i / 2i+1 of W are the value / gate of output column i; NW and n0 are even
            const float x = val(j), gt = val(j + 1);
            g.C[crow * g.ldc + (n >> 1)] = x * ((g.act == D4_ACT_GLU_SILU) ? siluf_(gt) : geluf_(gt));
        } else {
            float t = val(j);
            if (g.act == D4_ACT_SILU) t = siluf_(t);
            if (g.residual) t += g.residual[crow * g.ldr + n];
            g.C[crow * g.ldc + n] = t;
        }
    }
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

int d4_gemm_skinny_supported(const GemmArgs& g) {
    const bool glu = (g.act == D4_ACT_GLU_SILU || g.act == D4_ACT_GLU_GELU);
    return (g.M >= 1 && g.M <= 32 && g.N >= 1 && g.K >= 4 && (g.K % 4) == 0 && !g.transA && !g.transW && !g.ss_out && g.W && al16(g.A) && al16(g.W) &&
            (g.lda % 4) == 0 && (g.ldw % 4) == 0 && (((long long)g.amap.goff * g.lda) % 4) == 0 && (((long long)g.amap.gstride * g.lda) % 4) == 0 &&
            (!glu || (g.N % 2) == 0)) ? 1 : 0;
}

int d4_gemm_skinny(const GemmArgs& g, cudaStream_t stream) {
    if (!d4_gemm_skinny_supported(g)) return d4_fail("gemm_skinny: shape / alignment not supported (M <= 32, K %% 4 == 0, 16-byte aligned rows)");
#define D4_SK(MT, NW) gemm_skinny_kernel<MT, NW><<<(g.N + NW - 1) / NW, NTHR, 0, stream>>>(g)
    if (g.M <= 4) D4_SK(4, 4);
    else if (g.M <= 8) D4_SK(8, 4);
    else if (g.M <= 16) D4_SK(16, 4);
    else D4_SK(32, 2);
#undef D4_SK
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError());
    return 0;
}
