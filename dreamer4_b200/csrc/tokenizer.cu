// Front / back end kernels of the video tokenizer's per-frame pass (reference dreamer4.py:3833-3838, 3881-3886, 3569-3572,
// 4223-4227): patch gather, token assembly (LayerNorm of the patch projection + positional table + special tokens), tanh, and
// the un-patch + flow-matching Euler update.  All HBM-bound one-touch kernels; the projections between them are GEMMs
// (d4_linear_rows).  STATUS: drafted in round 1 after the GPU budget was spent - compiles for sm_100a, not yet run on hardware.
#include "kernels.h"

namespace {

// 'b c (h p1) (w p2) -> (b h w) (p1 p2 c)'; frame element (b, c, y, x) at frame[b * sb + c * sc + y * W + x]
__global__ void patchify_kernel(int B, int C, int H, int W, int p, const float* __restrict__ frame, long long sb, long long sc,
                                float* __restrict__ out) {
    const long long total = (long long)B * C * H * W;
    const int wp = W / p, hp = H / p, dp = p * p * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        // i enumerates OUTPUT elements: (row = (b, ph, pw), col = (p1, p2, c))
        const int col = (int)(i % dp); const long long row = i / dp;
        const int c = col % C, p2 = (col / C) % p, p1 = col / (C * p);
        const int pw = (int)(row % wp), ph = (int)((row / wp) % hp), b = (int)(row / ((long long)wp * hp));
        out[i] = frame[b * sb + c * sc + (long long)(ph * p + p1) * W + (pw * p + p2)];
    }
}

// frame <- frame + (pred - frame) * scale with pred = '(b h w) (p1 p2 c) -> b c (h p1) (w p2)' of the patch rows
__global__ void unpatchify_flow_kernel(int B, int C, int H, int W, int p, const float* __restrict__ patches, float* __restrict__ frame,
                                       long long sb, long long sc, float scale) {
    const long long total = (long long)B * C * H * W;
    const int wp = W / p, hp = H / p, dp = p * p * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        // i enumerates FRAME elements (b, c, y, x)
        const int x = (int)(i % W), y = (int)((i / W) % H), c = (int)((i / ((long long)W * H)) % C), b = (int)(i / ((long long)W * H * C));
        const long long row = ((long long)b * hp + y / p) * wp + x / p;
        const float pred = patches[row * dp + ((y % p) * p + (x % p)) * C + c];
        float* f = frame + b * sb + c * sc + (long long)y * W + x;
        const float v = *f;
        *f = v + (pred - v) * scale;
    }
}

// tokens (B, S, D): rows i < P = LayerNorm(lin[b * P + i]) * ln_w (+ pos_emb[i]); rows i >= P = special[b * special_bstride + (i - P) * D]
// nn.LayerNorm(bias=False), eps 1e-5, biased variance.  One warp per token row.
__global__ void tok_assemble_kernel(int B, int S, int P, int D, const float* __restrict__ lin, const float* __restrict__ ln_w,
                                    const float* __restrict__ pos_emb, const float* __restrict__ special, long long special_bstride,
                                    float* __restrict__ tokens) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= (long long)B * S) return;
    const int b = (int)(row / S), i = (int)(row % S);
    float* o = tokens + row * D;
    if (i >= P) {
        const float* sp = special + b * special_bstride + (long long)(i - P) * D;
        for (int c = lane; c < D; c += 32) o[c] = sp[c];
        return;
    }
    const float* x = lin + ((long long)b * P + i) * D;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) s += x[c];
    const float mean = warp_sum(s) / (float)D;
    float v = 0.f;
    for (int c = lane; c < D; c += 32) { const float t = x[c] - mean; v += t * t; }
    const float rstd = rsqrtf(warp_sum(v) / (float)D + D4_LN_EPS);
    for (int c = lane; c < D; c += 32) {
        float y = (x[c] - mean) * rstd * ln_w[c];
        if (pos_emb) y += pos_emb[(long long)i * D + c];
        o[c] = y;
    }
}

__global__ void tanh_kernel(float* __restrict__ x, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] = tanhf(x[i]);
}

inline unsigned grid_for(long long n, int threads) {
    const long long blocks = (n + threads - 1) / threads;
    return (unsigned)(blocks < 1 ? 1 : (blocks > 148LL * 16 ? 148LL * 16 : blocks));      // grid-stride beyond 16 CTAs per SM
}

}  // namespace

int d4_patchify_launch(int B, int C, int H, int W, int p, const float* frame, long long sb, long long sc, float* out, cudaStream_t s) {
    if (B < 1 || C < 1 || p < 1 || H % p || W % p) return d4_fail("d4_patchify: %d x %d image is not a multiple of patch size %d", H, W, p);
    patchify_kernel<<<grid_for((long long)B * C * H * W, 256), 256, 0, s>>>(B, C, H, W, p, frame, sb, sc, out);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError()); return 0;
}
int d4_unpatchify_flow_launch(int B, int C, int H, int W, int p, const float* patches, float* frame, long long sb, long long sc, float scale,
                              cudaStream_t s) {
    if (B < 1 || C < 1 || p < 1 || H % p || W % p) return d4_fail("d4_unpatchify_flow: %d x %d image is not a multiple of patch size %d", H, W, p);
    unpatchify_flow_kernel<<<grid_for((long long)B * C * H * W, 256), 256, 0, s>>>(B, C, H, W, p, patches, frame, sb, sc, scale);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError()); return 0;
}
int d4_tok_assemble_launch(int B, int S, int P, int D, const float* lin, const float* ln_w, const float* pos_emb, const float* special,
                           long long special_bstride, float* tokens, cudaStream_t s) {
    if (B < 1 || P < 0 || P > S || D < 1) return d4_fail("d4_tok_assemble: bad shape");
    const long long rows = (long long)B * S;
    tok_assemble_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(B, S, P, D, lin, ln_w, pos_emb, special, special_bstride, tokens);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError()); return 0;
}
int d4_tanh_launch(float* x, long long n, cudaStream_t s) {
    if (n <= 0) return 0;
    tanh_kernel<<<grid_for(n, 256), 256, 0, s>>>(x, n);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError()); return 0;
}
