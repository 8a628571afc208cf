// Exact-fp32 SIMT GEMM with the fused epilogues of the imagination pass.
// This is the "fp32" precision mode of the engine (bit-for-bit fp32 FMA accumulation, used by the
// parity tests and for the policy/value path where sampled action indices must match the oracle);
// the throughput path is the tcgen05 kernel in gemm_tc.cu with the same GemmArgs contract.
// transA / transW select (K,M) / (K,N) operand storage for the learn_from_experience backward GEMMs
// (dX = dY @ W needs W as (K,N); dW = dY^T @ X needs both operands reduced over their row index).
#include "kernels.h"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

// loads a BK x 64 tile of an operand into smem as T[k][row]
//   TRANS=false: operand is (rows, K) with leading dim ld, element (r,k) at P[map(r)*ld + k]
//   TRANS=true : operand is (K, rows), element (r,k) at P[k*ld + r]
template <bool VEC, bool TRANS>
__device__ __forceinline__ void load_tile(float (*T)[BM + 4], const float* __restrict__ P, long long ld, int r0, int rows, int k0, int K,
                                          const RowMap& map, int tid) {
    if (!TRANS) {
        const int lr = tid >> 2, lk = (tid & 3) * 4;
        const int r = r0 + lr, k = k0 + lk;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (r < rows) {
            const float* p = P + map(r) * ld + k;
            if (VEC) { if (k < K) { const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; } }
            else {
#pragma unroll
                for (int q = 0; q < 4; ++q) if (k + q < K) v[q] = p[q];
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) T[lk + q][lr] = v[q];
    } else {
        const int lk = tid >> 4, lr = (tid & 15) * 4;
        const int k = k0 + lk, r = r0 + lr;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (k < K) {
            const float* p = P + (long long)k * ld + r;
            if (VEC) { if (r < rows) { const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; } }
            else {
#pragma unroll
                for (int q = 0; q < 4; ++q) if (r + q < rows) v[q] = p[q];
            }
        }
        *reinterpret_cast<float4*>(&T[lk][lr]) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

template <bool VEC, bool TA, bool TW>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(GemmArgs g) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Ws[BK][BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int ty = tid >> 4, tx = tid & 15;           // compute: 16x16 threads, 4x4 each
    const RowMap ident = {0, 0, 0};

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < g.K; k0 += BK) {
        load_tile<VEC, TA>(As, g.A, g.lda, m0, g.M, k0, g.K, g.amap, tid);
        load_tile<VEC, TW>(Ws, g.W, g.ldw, n0, g.N, k0, g.K, ident, tid);
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 wv = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
            const float aa[4] = {av.x, av.y, av.z, av.w};
            const float ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], ww[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= g.M) continue;
        const long long crow = g.cmap(m);
        const float rs = g.row_scale ? g.row_scale[m] : 1.f;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            float t = acc[i][j] * rs;
            if (g.bias && n < g.N) t += g.bias[n];
            v[j] = t;
        }
        if (g.act == D4_ACT_GLU_SILU || g.act == D4_ACT_GLU_GELU) {
#pragma unroll
            for (int j = 0; j < 4; j += 2) {
                const int n = n0 + tx * 4 + j;
                if (n + 1 < g.N) {
                    const float gate = (g.act == D4_ACT_GLU_SILU) ? siluf_(v[j + 1]) : geluf_(v[j + 1]);
                    g.C[crow * g.ldc + (n >> 1)] = v[j] * gate;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + tx * 4 + j;
                if (n >= g.N) continue;
                float t = v[j];
                if (g.act == D4_ACT_SILU) t = siluf_(t);
                if (g.residual) t += g.residual[crow * g.ldr + n];
                g.C[crow * g.ldc + n] = t;
            }
        }
    }
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// A handful of output columns over many rows (the policy unembedding: N = number of actions, K = 4 D): the tile kernel above puts 64 rows
// on a CTA - 32 CTAs and 156 us at 2048 rows.  Here a warp owns a row: coalesced 16-byte loads of the row, every lane keeps N partial
// sums over its K slice (the N weight rows come from L1 / L2), one shuffle tree per column.  Plain products only, exact fp32.
constexpr int RD_MAXN = 8, RD_ROWS = 8;
__global__ void __launch_bounds__(32 * RD_ROWS) gemm_rowdot_kernel(GemmArgs g) {
    const int m = blockIdx.x * RD_ROWS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (m >= g.M) return;
    const float* a = g.A + (long long)m * g.lda;
    float acc[RD_MAXN];
#pragma unroll
    for (int n = 0; n < RD_MAXN; ++n) acc[n] = 0.f;
    for (int k = lane * 4; k < g.K; k += 128) {
        const float4 av = *reinterpret_cast<const float4*>(a + k);
#pragma unroll
        for (int n = 0; n < RD_MAXN; ++n) {
            if (n < g.N) {
                const float4 wv = __ldg(reinterpret_cast<const float4*>(g.W + (long long)n * g.ldw + k));
                acc[n] = fmaf(av.x, wv.x, acc[n]); acc[n] = fmaf(av.y, wv.y, acc[n]); acc[n] = fmaf(av.z, wv.z, acc[n]); acc[n] = fmaf(av.w, wv.w, acc[n]);
            }
        }
    }
#pragma unroll
    for (int n = 0; n < RD_MAXN; ++n) {
        if (n < g.N) {
            const float v = warp_sum(acc[n]);
            if (lane == 0) g.C[(long long)m * g.ldc + n] = v;
        }
    }
}

}  // namespace

int d4_gemm_simt(const GemmArgs& g, cudaStream_t stream) {
    if (g.M <= 0 || g.N <= 0) return 0;
    if (g.rs_mode || g.ss_out) return d4_fail("gemm_simt: sum-of-squares row statistics are only implemented by the CTA-pair tensor-core kernel");
    if (g.N <= RD_MAXN && g.M >= 256 && g.K % 4 == 0 && !g.transA && !g.transW && !g.bias && !g.row_scale && !g.residual && g.act == D4_ACT_NONE &&
        g.amap.grp == 0 && g.cmap.grp == 0 && al16(g.A) && al16(g.W) && g.lda % 4 == 0 && g.ldw % 4 == 0) {
        gemm_rowdot_kernel<<<(g.M + RD_ROWS - 1) / RD_ROWS, 32 * RD_ROWS, 0, stream>>>(g);
        D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError());
        return 0;
    }
    dim3 grid((g.M + BM - 1) / BM, (g.N + BN - 1) / BN);
    // 128-bit loads need every row start 16B-aligned and the contiguous extent a multiple of 4
    const bool va = al16(g.A) && (g.lda % 4 == 0) && ((g.transA ? g.M : g.K) % 4 == 0);
    const bool vw = al16(g.W) && (g.ldw % 4 == 0) && ((g.transW ? g.N : g.K) % 4 == 0);
    const bool vec = va && vw;
#define D4_LAUNCH(V, TA, TW) gemm_simt_kernel<V, TA, TW><<<grid, NT, 0, stream>>>(g)
    if (vec) {
        if (!g.transA && !g.transW) D4_LAUNCH(true, false, false);
        else if (!g.transA && g.transW) D4_LAUNCH(true, false, true);
        else if (g.transA && !g.transW) D4_LAUNCH(true, true, false);
        else D4_LAUNCH(true, true, true);
    } else {
        if (!g.transA && !g.transW) D4_LAUNCH(false, false, false);
        else if (!g.transA && g.transW) D4_LAUNCH(false, false, true);
        else if (g.transA && !g.transW) D4_LAUNCH(false, true, false);
        else D4_LAUNCH(false, true, true);
    }
#undef D4_LAUNCH
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError());
    return 0;
}
