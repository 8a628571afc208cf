// Row-wise warp-shuffle kernels of the imagination pass: RMS statistics, RMSNorm, LayerNorm+SiLU,
// token assembly, flow ODE step, HL-Gauss decode, gumbel-argmax action sampling, terminal bookkeeping.
// One warp per row; 128-bit loads where the row is 16-byte aligned.  All HBM-bound.
#include "kernels.h"

namespace {

constexpr int ROWS_PER_BLOCK = 8;   // 8 warps / CTA

__device__ __forceinline__ float row_sumsq(const float* __restrict__ x, int D, int lane) {
    float s = 0.f;
    if ((D & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) & 15) == 0)) {
        const float4* x4 = reinterpret_cast<const float4*>(x);
        for (int i = lane; i < (D >> 2); i += 32) { float4 t = x4[i]; s += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w; }
    } else {
        for (int i = lane; i < D; i += 32) { float t = x[i]; s += t * t; }
    }
    return warp_sum(s);
}

// rstd[m] = rsqrt(mean(x[m]^2) + eps)   (the RMSNorm statistic; gamma is folded into the next GEMM's weight)
__global__ void row_rstd_kernel(const float* __restrict__ x, long long ldx, RowMap map, int M, int D, float* __restrict__ out) {
    const int m = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (m >= M) return;
    const float ss = row_sumsq(x + map(m) * ldx, D, lane);
    if (lane == 0) out[m] = rsqrtf(ss / (float)D + D4_RMS_EPS);
}

// RMS statistic of the rows map(0..M-1) of x: rstd compactly (out_c[m]) and, for the consumers that index by the full row (the pool
// context gather), out_f[map(m)] = the sum of squares (full_is_ss) or the rstd
__global__ void row_stat_map_kernel(const float* __restrict__ x, long long ldx, RowMap map, int M, int D, float* __restrict__ out_c,
                                    float* __restrict__ out_f, int full_is_ss) {
    const int m = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (m >= M) return;
    const long long r = map(m);
    const float ss = row_sumsq(x + r * ldx, D, lane);
    if (lane == 0) {
        const float rstd = rsqrtf(ss / (float)D + D4_RMS_EPS);
        if (out_c) out_c[m] = rstd;
        if (out_f) out_f[r] = full_is_ss ? ss : rstd;
    }
}

__global__ void row_sumsq_kernel(const float* __restrict__ x, long long ldx, int M, int D, float* __restrict__ out) {
    const int m = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (m >= M) return;
    const float ss = row_sumsq(x + (long long)m * ldx, D, lane);
    if (lane == 0) out[m] = ss;
}

// out[m] = x[map(m)] * rstd * w        nn.RMSNorm (reference dreamer4.py:1906, 2089, 2822, 4831)
__global__ void rmsnorm_rows_kernel(const float* __restrict__ x, long long ldx, RowMap map, const float* __restrict__ w,
                                    int M, int D, float* __restrict__ out, long long ldo) {
    const int m = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (m >= M) return;
    const float* xr = x + map(m) * ldx;
    const float r = rsqrtf(row_sumsq(xr, D, lane) / (float)D + D4_RMS_EPS);
    float* o = out + (long long)m * ldo;
    for (int i = lane; i < D; i += 32) o[i] = xr[i] * r * w[i];
}

// y = act(LayerNorm(x) * w + b); optionally stores (mean, rstd) for the backward pass.
// x-mlps normed-MLP hidden layer: Linear -> LayerNorm -> SiLU (oracle/shims/x_mlps_pytorch/normed_mlp.py).
__global__ void ln_act_rows_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ w, const float* __restrict__ b,
                                   int M, int D, float* __restrict__ out, long long ldo, int act,
                                   float* __restrict__ save_mean, float* __restrict__ save_rstd) {
    const int m = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (m >= M) return;
    const float* xr = x + (long long)m * ldx;
    float s = 0.f;
    for (int i = lane; i < D; i += 32) s += xr[i];
    const float mean = warp_sum(s) / (float)D;
    float v = 0.f;
    for (int i = lane; i < D; i += 32) { float t = xr[i] - mean; v += t * t; }
    const float rstd = rsqrtf(warp_sum(v) / (float)D + D4_LN_EPS);
    float* o = out + (long long)m * ldo;
    for (int i = lane; i < D; i += 32) {
        float t = (xr[i] - mean) * rstd * w[i] + b[i];
        o[i] = (act == D4_ACT_SILU) ? siluf_(t) : t;
    }
    if (lane == 0 && save_mean) { save_mean[m] = mean; save_rstd[m] = rstd; }
}

// The same with the row held in registers (D <= 128 * VPL, 16-byte aligned rows): one pass over memory, VPL independent 16-byte loads
// per lane in flight instead of three dependent scalar sweeps - the head MLPs' 2048-wide rows took 39 us per launch at 2048 rows
// against ~6 us of traffic.
template <int VPL>
__global__ void __launch_bounds__(32 * ROWS_PER_BLOCK) ln_act_rows_reg_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ w,
                                                                              const float* __restrict__ b, int M, int D, float* __restrict__ out, long long ldo,
                                                                              int act, float* __restrict__ save_mean, float* __restrict__ save_rstd) {
    const int m = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (m >= M) return;
    const float4* xr = reinterpret_cast<const float4*>(x + (long long)m * ldx);
    const int n4 = D >> 2;
    float4 v[VPL];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
        const int i = lane + 32 * k;
        v[k] = (i < n4) ? xr[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    }
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
        if (lane + 32 * k < n4) {
            const float tx = v[k].x - mean, ty = v[k].y - mean, tz = v[k].z - mean, tw = v[k].w - mean;
            q += (tx * tx + ty * ty) + (tz * tz + tw * tw);
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)D + D4_LN_EPS);
    float4* o = reinterpret_cast<float4*>(out + (long long)m * ldo);
    const float4* w4 = reinterpret_cast<const float4*>(w);
    const float4* b4 = reinterpret_cast<const float4*>(b);
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
        const int i = lane + 32 * k;
        if (i < n4) {
            const float4 ww = __ldg(w4 + i), bb = __ldg(b4 + i);
            float4 t;
            t.x = (v[k].x - mean) * rstd * ww.x + bb.x; t.y = (v[k].y - mean) * rstd * ww.y + bb.y;
            t.z = (v[k].z - mean) * rstd * ww.z + bb.z; t.w = (v[k].w - mean) * rstd * ww.w + bb.w;
            if (act == D4_ACT_SILU) { t.x = siluf_(t.x); t.y = siluf_(t.y); t.z = siluf_(t.z); t.w = siluf_(t.w); }
            o[i] = t;
        }
    }
    if (lane == 0 && save_mean) { save_mean[m] = mean; save_rstd[m] = rstd; }
}

// Non-latent tokens of the newest frame (reference dreamer4.py:7004-7010, 7101-7126, 7193-7222):
//   s = 0                      flow token  = cat(signal_levels_embed[signal], step_size_embed[step])
//   s = 1 .. nsp               spatial tokens (written by the latents->spatial GEMM, not here)
//   s = 1+nsp .. 1+nsp+nreg-1  register tokens
//   s = S-2 (if has_actions)   action token = 0 at frame 0, else action_learned_embed + sum_types embed[a + offset]
//   s = S-1                    agent token  = agent_learned_embed (+ task_embed[task])
__global__ void assemble_tokens_kernel(AssembleArgs a) {
    const int b = blockIdx.x;
    const int nfix = 1 + a.nreg + a.has_actions + 1;
    for (int f = threadIdx.x >> 5; f < nfix; f += (blockDim.x >> 5)) {
        const int lane = threadIdx.x & 31;
        int s; int kind;   // 0 flow, 1 register, 2 action, 3 agent
        if (f == 0) { s = 0; kind = 0; }
        else if (f <= a.nreg) { s = a.nsp + f; kind = 1; }
        else if (a.has_actions && f == a.nreg + 1) { s = a.S - 2; kind = 2; }
        else { s = a.S - 1; kind = 3; }
        float* o = a.tokens + ((long long)b * a.S + s) * a.D;
        const int half = a.D >> 1;
        const long long sig = a.signal_rows ? a.signal_rows[b] : a.signal, stp = a.step_rows ? a.step_rows[b] : a.step;
        for (int i = lane; i < a.D; i += 32) {
            float v;
            if (kind == 0) v = (i < half) ? a.sig_emb[sig * half + i] : a.step_emb[stp * half + (i - half)];
            else if (kind == 1) v = a.registers[(long long)(f - 1) * a.D + i];
            else if (kind == 2) {
                if (a.prev_actions == nullptr) v = 0.f;
                else {
                    float e = 0.f;
                    for (int t = 0; t < a.na; ++t) {
                        const long long id = a.prev_actions[(long long)b * a.pa_stride + t] + a.act_off[t];
                        e += a.action_emb[id * a.D + i];
                    }
                    v = a.action_learned[i] + e;
                }
            } else {
                v = a.agent_embed[i];
                if (a.tasks) v += a.task_emb[a.tasks[b] * a.D + i];
            }
            o[i] = v;
        }
    }
}

// x-space shortcut flow Euler step (reference dreamer4.py:6567-6580, times 5408-5419):
//   x += (pred - x) / (1 - tau) * dt
__global__ void flow_step_kernel(float* __restrict__ x, const float* __restrict__ pred, long long n, float one_minus_tau, float dt) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const float xv = x[i]; x[i] = xv + (pred[i] - xv) / one_minus_tau * dt; }
}

// out[b, n, :] (row stride given) = clamp(x[b, n, :], -1, 1)   (reference dreamer4.py:6686)
__global__ void store_latents_kernel(const float* __restrict__ x, float* __restrict__ out, int B, long long per_b, long long out_bstride) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * per_b) return;
    const long long b = i / per_b, r = i % per_b;
    out[b * out_bstride + r] = fminf(fmaxf(x[i], -1.f), 1.f);
}

// strided row copy (agent embeds, logits into the (B,T,...) experience tensors)
__global__ void copy_rows_kernel(const float* __restrict__ src, long long lds, float* __restrict__ dst, long long ldd, int M, int D) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)M * D) return;
    const long long m = i / D, c = i % D;
    dst[m * ldd + c] = src[m * lds + c];
}

// dst[m, :] = src[map(m), :] (row gather through a grouped row map; D = 1 gathers a per-row statistic)
__global__ void gather_rows_kernel(const float* __restrict__ src, long long lds, RowMap map, long long M, int D, float* __restrict__ dst, long long ldd) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * D) return;
    const long long m = i / D, c = i % D;
    dst[m * ldd + c] = src[map((int)m) * lds + c];
}

// HL-Gauss expectation: softmax(logits) . centers   (reference dreamer4.py:1094-1105 via hl_gauss transform_from_logits)
__global__ void hl_gauss_decode_kernel(const float* __restrict__ logits, long long ld, int M, int K, const float* __restrict__ centers,
                                       float* __restrict__ out, long long out_stride) {
    const int m = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (m >= M) return;
    const float* l = logits + (long long)m * ld;
    float mx = -INFINITY;
    for (int i = lane; i < K; i += 32) mx = fmaxf(mx, l[i]);
    mx = warp_max(mx);
    float se = 0.f, sc = 0.f;
    for (int i = lane; i < K; i += 32) { const float e = expf(l[i] - mx); se += e; sc += e * centers[i]; }
    se = warp_sum(se); sc = warp_sum(sc);
    if (lane == 0) out[(long long)m * out_stride] = sc / se;
}

// Gumbel-argmax sampling + log-prob per action type (MultiCategorical.sample / log_prob, reference call sites
// dreamer4.py:1375-1376, 1422-1426, 6628-6657; gumbel form dreamer4.py:473-497):
//   idx = argmax(logits / max(T, 1e-10) - log(-log(u))),  log(t) = log(max(t, 1e-20));  logp = log_softmax(logits)[idx]
// one warp per (b, action type); ties resolve to the lowest index like torch.argmax.
__global__ void sample_actions_kernel(const float* __restrict__ logits, long long ld, const float* __restrict__ u, long long ldu,
                                      int B, int na, const int* __restrict__ sizes_offs /* na sizes then na offsets */,
                                      float inv_temp, long long* __restrict__ actions, long long act_stride,
                                      float* __restrict__ logp, long long lp_stride) {
    const int w = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (w >= B * na) return;
    const int b = w / na, t = w % na;
    const int n = sizes_offs[t], off = sizes_offs[na + t];
    const float* l = logits + (long long)b * ld + off;
    const float* uu = u + (long long)b * ldu + off;
    float best = -INFINITY; int besti = 0x7fffffff;
    float mx = -INFINITY;
    for (int i = lane; i < n; i += 32) {
        const float li = l[i];
        const float g = -logf(fmaxf(-logf(fmaxf(uu[i], 1e-20f)), 1e-20f));
        const float v = li * inv_temp + g;
        if (v > best) { best = v; besti = i; }
        mx = fmaxf(mx, li);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(D4_FULL, best, o);
        const int oi = __shfl_xor_sync(D4_FULL, besti, o);
        if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
    }
    mx = warp_max(mx);
    float se = 0.f;
    for (int i = lane; i < n; i += 32) se += expf(l[i] - mx);
    se = warp_sum(se);
    if (lane == 0) {
        actions[(long long)b * act_stride + t] = besti;
        logp[(long long)b * lp_stride + t] = l[besti] - mx - logf(se);
    }
}

// mean over the N latent tokens of each sample (terminal head input, reference dreamer4.py:6605-6607)
__global__ void mean_tokens_kernel(const float* __restrict__ x, int B, int N, int Dl, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * Dl) return;
    const int b = i / Dl, c = i % Dl;
    float s = 0.f;
    for (int n = 0; n < N; ++n) s += x[((long long)b * N + n) * Dl + c];
    out[i] = s / (float)N;
}

// Bernoulli terminal draw + first-termination bookkeeping (reference dreamer4.py:6608-6616):
//   is_term = u < sigmoid(logit); lens[b] = frame+1 where newly terminated; terminals |= is_term
__global__ void terminal_update_kernel(const float* __restrict__ logit, long long ld, const float* __restrict__ u, int B, int frame,
                                       long long* __restrict__ lens, unsigned char* __restrict__ terminals) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const bool is_term = u[b] < sigmoidf_(logit[(long long)b * ld]);
    const bool was = terminals[b] != 0;
    if (is_term && !was) lens[b] = frame + 1;
    terminals[b] = (was || is_term) ? 1 : 0;
}

inline int nblk(long long n, int per) { return (int)((n + per - 1) / per); }

}  // namespace

int d4_row_rstd(const float* x, long long ldx, RowMap map, int M, int D, float* out, cudaStream_t s) {
    if (M <= 0) return 0;
    row_rstd_kernel<<<nblk(M, ROWS_PER_BLOCK), 32 * ROWS_PER_BLOCK, 0, s>>>(x, ldx, map, M, D, out);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError()); return 0;
}
int d4_row_stat_map(const float* x, long long ldx, RowMap map, int M, int D, float* out_c, float* out_f, int full_is_ss, cudaStream_t s) {
    if (M <= 0) return 0;
    row_stat_map_kernel<<<nblk(M, ROWS_PER_BLOCK), 32 * ROWS_PER_BLOCK, 0, s>>>(x, ldx, map, M, D, out_c, out_f, full_is_ss);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError()); return 0;
}
int d4_row_sumsq(const float* x, long long ldx, int M, int D, float* out, cudaStream_t s) {
    if (M <= 0) return 0;
    row_sumsq_kernel<<<nblk(M, ROWS_PER_BLOCK), 32 * ROWS_PER_BLOCK, 0, s>>>(x, ldx, M, D, out);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError()); return 0;
}
int d4_rmsnorm_rows(const float* x, long long ldx, RowMap map, const float* w, int M, int D, float* out, long long ldo, cudaStream_t s) {
    if (M <= 0) return 0;
    rmsnorm_rows_kernel<<<nblk(M, ROWS_PER_BLOCK), 32 * ROWS_PER_BLOCK, 0, s>>>(x, ldx, map, w, M, D, out, ldo);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError()); return 0;
}
int d4_ln_act_rows(const float* x, long long ldx, const float* w, const float* b, int M, int D, float* out, long long ldo, int act,
                   float* save_mean, float* save_rstd, cudaStream_t s) {
    if (M <= 0) return 0;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (D % 4 == 0 && D <= 2048 && ldx % 4 == 0 && ldo % 4 == 0 && al16(x) && al16(out) && al16(w) && al16(b)) {
        const unsigned grid = (unsigned)nblk(M, ROWS_PER_BLOCK);
        if (D <= 512)       ln_act_rows_reg_kernel<4><<<grid, 32 * ROWS_PER_BLOCK, 0, s>>>(x, ldx, w, b, M, D, out, ldo, act, save_mean, save_rstd);
        else if (D <= 1024) ln_act_rows_reg_kernel<8><<<grid, 32 * ROWS_PER_BLOCK, 0, s>>>(x, ldx, w, b, M, D, out, ldo, act, save_mean, save_rstd);
        else                ln_act_rows_reg_kernel<16><<<grid, 32 * ROWS_PER_BLOCK, 0, s>>>(x, ldx, w, b, M, D, out, ldo, act, save_mean, save_rstd);
        D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError()); return 0;
    }
    ln_act_rows_kernel<<<nblk(M, ROWS_PER_BLOCK), 32 * ROWS_PER_BLOCK, 0, s>>>(x, ldx, w, b, M, D, out, ldo, act, save_mean, save_rstd);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError()); return 0;
}
int d4_assemble_tokens(const AssembleArgs& a, cudaStream_t s) {
    assemble_tokens_kernel<<<a.B, 128, 0, s>>>(a);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError()); return 0;
}
int d4_flow_step(float* x, const float* pred, long long n, float one_minus_tau, float dt, cudaStream_t s) {
    flow_step_kernel<<<nblk(n, 256), 256, 0, s>>>(x, pred, n, one_minus_tau, dt);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError()); return 0;
}
int d4_store_latents(const float* x, float* out, int B, long long per_b, long long out_bstride, cudaStream_t s) {
    store_latents_kernel<<<nblk((long long)B * per_b, 256), 256, 0, s>>>(x, out, B, per_b, out_bstride);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError()); return 0;
}
int d4_copy_rows(const float* src, long long lds, float* dst, long long ldd, int M, int D, cudaStream_t s) {
    if (M <= 0 || D <= 0) return 0;
    copy_rows_kernel<<<nblk((long long)M * D, 256), 256, 0, s>>>(src, lds, dst, ldd, M, D);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError()); return 0;
}
int d4_gather_rows(const float* src, long long lds, RowMap map, long long M, int D, float* dst, long long ldd, cudaStream_t s) {
    if (M <= 0 || D <= 0) return 0;
    gather_rows_kernel<<<nblk(M * D, 256), 256, 0, s>>>(src, lds, map, M, D, dst, ldd);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError()); return 0;
}
int d4_hl_gauss_decode(const float* logits, long long ld, int M, int K, const float* centers, float* out, long long out_stride, cudaStream_t s) {
    if (M <= 0) return 0;
    hl_gauss_decode_kernel<<<nblk(M, ROWS_PER_BLOCK), 32 * ROWS_PER_BLOCK, 0, s>>>(logits, ld, M, K, centers, out, out_stride);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError()); return 0;
}
int d4_sample_actions(const float* logits, long long ld, const float* u, long long ldu, int B, int na, const int* sizes_offs, float inv_temp,
                      long long* actions, long long act_stride, float* logp, long long lp_stride, cudaStream_t s) {
    if (B <= 0 || na <= 0) return 0;
    sample_actions_kernel<<<nblk((long long)B * na, ROWS_PER_BLOCK), 32 * ROWS_PER_BLOCK, 0, s>>>(logits, ld, u, ldu, B, na, sizes_offs, inv_temp,
                                                                                                 actions, act_stride, logp, lp_stride);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError()); return 0;
}
int d4_mean_tokens(const float* x, int B, int N, int Dl, float* out, cudaStream_t s) {
    mean_tokens_kernel<<<nblk((long long)B * Dl, 256), 256, 0, s>>>(x, B, N, Dl, out);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError()); return 0;
}
int d4_terminal_update(const float* logit, long long ld, const float* u, int B, int frame, long long* lens, unsigned char* terminals, cudaStream_t s) {
    terminal_update_kernel<<<nblk(B, 256), 256, 0, s>>>(logit, ld, u, B, frame, lens, terminals);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError()); return 0;
}
