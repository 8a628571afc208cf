// learn_from_experience (reference dreamer4.py:5893-6305, objective 'ppo', only_learn_policy_value_heads=True):
// lambda-return warp scan (calc_gae, 1566-1600), advantage z-score (404-410, 6017-6024), policy/value MLP forward,
// PPO clip surrogate x delight gate + entropy bonus (6111-6242), HL-Gauss soft cross-entropy (6254-6295) and the
// full backward of both heads, producing losses and parameter gradients in one call.  Exact fp32 throughout
// (the north star asks for losses within 1e-4 relative of the reference).
#include <string.h>
#include <algorithm>
#include "engine.h"

#define D4_TRY(expr) do { int rc__ = (expr); if (rc__ != 0) return rc__; } while (0)

namespace {

constexpr int RPB = 8;   // rows (warps) per block

// ---- calc_gae: one warp per batch row, Kogge-Stone scan of the affine maps out -> delta + gate*out, right to left
__global__ void gae_kernel(int B, int T, const float* __restrict__ rewards, const float* __restrict__ values,
                           const unsigned char* __restrict__ masks, const unsigned char* __restrict__ learn_masks,
                           float gamma, float lam, float* __restrict__ returns) {
    const int b = blockIdx.x * RPB + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B) return;
    const float* r = rewards + (long long)b * T; const float* v = values + (long long)b * T;
    const unsigned char* mk = masks + (long long)b * T; const unsigned char* lm = learn_masks + (long long)b * T;
    float carry = 0.f;
    for (int base = 0; base < T; base += 32) {
        const int i = base + lane;            // position in reversed order
        const int t = T - 1 - i;
        float a = 0.f, d = 0.f, vt = 0.f;
        if (i < T) {
            const float m = mk[t] ? 1.f : 0.f;
            vt = v[t];
            const float vnext = (t + 1 < T) ? v[t + 1] : 0.f;
            d = r[t] + gamma * vnext * m - vt;
            if (!lm[t]) d = 0.f;
            a = gamma * lam * m;
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float ap = __shfl_up_sync(D4_FULL, a, o), dp = __shfl_up_sync(D4_FULL, d, o);
            if (lane >= o) { d = d + a * dp; a = a * ap; }
        }
        const float out = d + a * carry;
        if (i < T) returns[(long long)b * T + t] = out + vt;
        carry = __shfl_sync(D4_FULL, out, 31);
    }
}

// ---- masks + masked rewards/values (reference dreamer4.py:5943-5967)
__global__ void learn_masks_kernel(int B, int T, const float* __restrict__ rewards, const float* __restrict__ values,
                                   const long long* __restrict__ lens, const unsigned char* __restrict__ is_truncated,
                                   float* __restrict__ r_out, float* __restrict__ v_out, unsigned char* __restrict__ gae_mask,
                                   unsigned char* __restrict__ learn_mask) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * T) return;
    const int b = (int)(i / T), t = (int)(i % T);
    const long long len = lens[b];
    const bool in = t < len;
    r_out[i] = in ? rewards[i] : 0.f;
    v_out[i] = in ? values[i] : 0.f;
    learn_mask[i] = t < (len - (is_truncated[b] ? 1 : 0));
    const long long lm1 = len - 1 > 0 ? len - 1 : 0;
    gae_mask[i] = t < lm1;
}

__device__ float block_sum(float v, float* sh) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    float t = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.f;
    if (w == 0) t = warp_sum(t);
    if (threadIdx.x == 0) sh[0] = t;
    __syncthreads();
    const float r = sh[0];
    __syncthreads();
    return r;
}

// advantage = returns - old_values, z-scored over the learnable mask (population variance, eps clamp);
// stats[0] = number of learnable steps
// ema (keep_reward_ema_stats, 5987-6013): [mean, std] of the running return statistics - both terms are normalised before the
// subtraction, in the reference's order
__global__ void adv_stats_kernel(long long R, const float* __restrict__ returns, const float* __restrict__ values,
                                 const unsigned char* __restrict__ mask, int normalize, float eps, const float* __restrict__ ema,
                                 float* __restrict__ adv, float* __restrict__ stats) {
    __shared__ float sh[32];
    float n = 0.f, s = 0.f;
    const float em = ema ? ema[0] : 0.f, es = ema ? ema[1] : 1.f;
    for (long long i = threadIdx.x; i < R; i += blockDim.x) {
        const float a = ema ? (returns[i] - em) / es - (values[i] - em) / es : returns[i] - values[i];
        adv[i] = a;
        if (mask[i]) { n += 1.f; s += a; }
    }
    n = block_sum(n, sh); s = block_sum(s, sh);
    const float mean = n > 0.f ? s / n : 0.f;
    float q = 0.f;
    for (long long i = threadIdx.x; i < R; i += blockDim.x) if (mask[i]) { const float dlt = adv[i] - mean; q += dlt * dlt; }
    q = block_sum(q, sh);
    const float var = n > 0.f ? q / n : 0.f;
    if (normalize) {
        const float den = sqrtf(fmaxf(var, eps));
        for (long long i = threadIdx.x; i < R; i += blockDim.x) adv[i] = (adv[i] - mean) / den;
    }
    if (threadIdx.x == 0) { stats[0] = n; stats[1] = mean; stats[2] = var; }
}

// ---- policy row kernel: per (b,t) row -> surrogate / entropy loss terms and d(total_policy_loss)/d(logits)
// objective: D4_OBJECTIVE_PPO (clipped ratio, 6200-6212), _SPO (quadratic trust region, 6184-6198), _PMPO (sign-weighted
// log-likelihood + KL to the stored unembeds, 6127-6182)
struct PpoArgs {
    int R, na, A_total; const int* sizes_offs;
    const float* logits; long long ld;
    const long long* actions; const float* old_logp; const float* adv; const unsigned char* mask; const float* stats;
    float eps_clip, entropy_weight, delight_temp; int use_gate;
    int objective; float pmpo_alpha, pmpo_kl_weight; int pmpo_reverse_kl; const float* old_logits; long long ld_old;
    float* dlogits; float* row_pl; float* row_ent;
};
__global__ void ppo_row_kernel(PpoArgs p) {
    const int r = blockIdx.x * RPB + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= p.R) return;
    const float* l = p.logits + (long long)r * p.ld;
    float* dl = p.dlogits + (long long)r * p.ld;
    const bool on = p.mask[r] != 0;
    const float n = p.stats[0];
    const float inv_n = (on && n > 0.f) ? 1.f / n : 0.f;
    float logp = 0.f, oldlp = 0.f, ent_sum = 0.f;
    // pass 1: per action type log-softmax statistics
    float mx_t[D4_MAX_ACTION_TYPES], lse_t[D4_MAX_ACTION_TYPES], H_t[D4_MAX_ACTION_TYPES];
    for (int t = 0; t < p.na; ++t) {
        const int sz = p.sizes_offs[t], off = p.sizes_offs[p.na + t];
        float mx = -INFINITY;
        for (int i = lane; i < sz; i += 32) mx = fmaxf(mx, l[off + i]);
        mx = warp_max(mx);
        float se = 0.f;
        for (int i = lane; i < sz; i += 32) se += expf(l[off + i] - mx);
        se = warp_sum(se);
        const float lse = logf(se);
        float h = 0.f;
        for (int i = lane; i < sz; i += 32) { const float lp = l[off + i] - mx - lse; h -= expf(lp) * lp; }
        h = warp_sum(h);
        mx_t[t] = mx; lse_t[t] = lse; H_t[t] = h;
        const long long a = p.actions[(long long)r * p.na + t];
        logp += l[off + a] - mx - lse;
        oldlp += p.old_logp[(long long)r * p.na + t];
        ent_sum += h;
    }
    const float A = p.adv[r];
    const float gate = p.use_gate ? sigmoidf_((-logp * A) / p.delight_temp) : 1.f;      // detached in the reference (6119-6120)
    float pl, dlogp;
    if (p.objective == D4_OBJECTIVE_PMPO) {
        // -alpha * (sum_{A>=0} g*logp*|tanh A| - sum_{A<0} g*logp*|tanh A|) / n
        const float w = p.pmpo_alpha * fabsf(tanhf(A)) * gate * (A >= 0.f ? 1.f : -1.f);
        pl = -w * logp;
        dlogp = -w * inv_n;
    } else {
        const float ratio = expf(logp - oldlp);
        if (p.objective == D4_OBJECTIVE_SPO) {
            const float q = fabsf(A) / (2.f * p.eps_clip), d = ratio - 1.f;
            pl = -(ratio * A - q * d * d) * gate;
            dlogp = -(A - 2.f * q * d) * ratio * gate * inv_n;
        } else {
            const float clipped = fminf(fmaxf(ratio, 1.f - p.eps_clip), 1.f + p.eps_clip);
            const float s1 = ratio * A, s2 = clipped * A;
            pl = -fminf(s1, s2) * gate;
            // d(-min(s1,s2))/dlogp: the unclipped branch (or the tie inside the clip range) passes A*ratio, the clipped branch 0
            dlogp = (s1 <= s2) ? -gate * A * ratio * inv_n : 0.f;
        }
    }
    // pass 2: gradients
    for (int t = 0; t < p.na; ++t) {
        const int sz = p.sizes_offs[t], off = p.sizes_offs[p.na + t];
        const long long a = p.actions[(long long)r * p.na + t];
        for (int i = lane; i < sz; i += 32) {
            const float lp = l[off + i] - mx_t[t] - lse_t[t];
            const float pj = expf(lp);
            float gl = dlogp * (((long long)i == a ? 1.f : 0.f) - pj);
            gl += p.entropy_weight * inv_n * pj * (lp + H_t[t]);
            dl[off + i] = gl;
        }
    }
    if (p.objective == D4_OBJECTIVE_PMPO && p.pmpo_kl_weight > 0.f) {
        // KL between the stored and the replayed unembeds, taken as the reference takes it: ONE categorical over the
        // flat (A_total) logits (6160-6169 pass the unsplit tensors to kl_div)
        const float* o = p.old_logits + (long long)r * p.ld_old;
        float mxn = -INFINITY, mxo = -INFINITY;
        for (int i = lane; i < p.A_total; i += 32) { mxn = fmaxf(mxn, l[i]); mxo = fmaxf(mxo, o[i]); }
        mxn = warp_max(mxn); mxo = warp_max(mxo);
        float sn = 0.f, so = 0.f;
        for (int i = lane; i < p.A_total; i += 32) { sn += expf(l[i] - mxn); so += expf(o[i] - mxo); }
        const float lsn = mxn + logf(warp_sum(sn)), lso = mxo + logf(warp_sum(so));
        float kl = 0.f;
        for (int i = lane; i < p.A_total; i += 32) {
            const float ln = l[i] - lsn, lo = o[i] - lso;
            kl += p.pmpo_reverse_kl ? expf(lo) * (lo - ln) : expf(ln) * (ln - lo);
        }
        kl = warp_sum(kl);
        const float wk = p.pmpo_kl_weight * inv_n;
        __syncwarp();       // pass 2 stored dl[] per action type: a different lane owns each flat index here
        for (int i = lane; i < p.A_total; i += 32) {
            const float ln = l[i] - lsn, lo = o[i] - lso, pn = expf(ln);
            dl[i] += wk * (p.pmpo_reverse_kl ? pn - expf(lo) : pn * (ln - lo - kl));
        }
        pl += p.pmpo_kl_weight * kl;
    }
    if (lane == 0) { p.row_pl[r] = on ? pl : 0.f; p.row_ent[r] = on ? -ent_sum : 0.f; }
}

// ---- value row kernel: soft cross-entropy against HL-Gauss(returns) (reference dreamer4.py:6277-6295; hl_gauss to_probs)
struct ValueArgs {
    int R, K; const float* bins; long long ld; const float* returns; const unsigned char* mask; const float* stats;
    const float* support; float sigma_sqrt2, hl_eps, lo, hi;
    float* dbins; float* row_vl;
};
__global__ void value_row_kernel(ValueArgs p) {
    const int r = blockIdx.x * RPB + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= p.R) return;
    const float* l = p.bins + (long long)r * p.ld;
    float* dl = p.dbins + (long long)r * p.ld;
    const bool on = p.mask[r] != 0;
    const float n = p.stats[0];
    const float inv_n = (on && n > 0.f) ? 1.f / n : 0.f;
    const float target = fminf(fmaxf(p.returns[r], p.lo), p.hi);
    const float z = erff((p.support[p.K] - target) / p.sigma_sqrt2) - erff((p.support[0] - target) / p.sigma_sqrt2);
    const float zc = fmaxf(z, p.hl_eps);
    float mx = -INFINITY;
    for (int i = lane; i < p.K; i += 32) mx = fmaxf(mx, l[i]);
    mx = warp_max(mx);
    float se = 0.f;
    for (int i = lane; i < p.K; i += 32) se += expf(l[i] - mx);
    se = warp_sum(se);
    const float lse = logf(se);
    float loss = 0.f, psum = 0.f;
    for (int i = lane; i < p.K; i += 32) {
        const float pr = (erff((p.support[i + 1] - target) / p.sigma_sqrt2) - erff((p.support[i] - target) / p.sigma_sqrt2)) / zc;
        loss -= pr * (l[i] - mx - lse);
        psum += pr;
    }
    loss = warp_sum(loss); psum = warp_sum(psum);
    for (int i = lane; i < p.K; i += 32) {
        const float pr = (erff((p.support[i + 1] - target) / p.sigma_sqrt2) - erff((p.support[i] - target) / p.sigma_sqrt2)) / zc;
        dl[i] = inv_n * (expf(l[i] - mx - lse) * psum - pr);
    }
    if (lane == 0) p.row_vl[r] = on ? loss : 0.f;
}

// losses = [policy_total, value, surrogate, entropy_term]
__global__ void loss_reduce_kernel(long long R, const float* __restrict__ row_pl, const float* __restrict__ row_ent,
                                   const float* __restrict__ row_vl, const float* __restrict__ stats, float entropy_weight,
                                   float* __restrict__ losses) {
    __shared__ float sh[32];
    float a = 0.f, b = 0.f, c = 0.f;
    for (long long i = threadIdx.x; i < R; i += blockDim.x) { a += row_pl[i]; b += row_ent[i]; c += row_vl[i]; }
    a = block_sum(a, sh); b = block_sum(b, sh); c = block_sum(c, sh);
    if (threadIdx.x == 0) {
        const float n = stats[0];
        const float pl = n > 0.f ? a / n : 0.f, el = n > 0.f ? b / n : 0.f, vl = n > 0.f ? c / n : 0.f;
        losses[0] = pl + el * entropy_weight; losses[1] = vl; losses[2] = pl; losses[3] = el;
    }
}

// ---- backward of  x' = SiLU(LayerNorm(y) * w + b): dy (in place over dx), one warp per row
__global__ void ln_silu_bwd_rows_kernel(int M, int W, const float* __restrict__ y, const float* __restrict__ mean, const float* __restrict__ rstd,
                                        const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ dx) {
    const int m = blockIdx.x * RPB + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (m >= M) return;
    const float* yr = y + (long long)m * W; float* dr = dx + (long long)m * W;
    const float mu = mean[m], rs = rstd[m];
    float s1 = 0.f, s2 = 0.f;
    for (int i = lane; i < W; i += 32) {
        const float yh = (yr[i] - mu) * rs;
        const float z = yh * w[i] + b[i];
        const float sg = sigmoidf_(z);
        const float dz = dr[i] * (sg * (1.f + z * (1.f - sg)));
        const float dzg = dz * w[i];
        s1 += dzg; s2 += dzg * yh;
    }
    s1 = warp_sum(s1) / (float)W; s2 = warp_sum(s2) / (float)W;
    for (int i = lane; i < W; i += 32) {
        const float yh = (yr[i] - mu) * rs;
        const float z = yh * w[i] + b[i];
        const float sg = sigmoidf_(z);
        const float dz = dr[i] * (sg * (1.f + z * (1.f - sg)));
        dr[i] = rs * (dz * w[i] - s1 - yh * s2);
    }
}

// dlnw += sum_rows dz * yhat, dlnb += sum_rows dz   (dz recomputed from the upstream gradient; must run BEFORE the in-place row kernel)
__global__ void ln_param_grad_kernel(int M, int W, const float* __restrict__ y, const float* __restrict__ mean, const float* __restrict__ rstd,
                                     const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ dx,
                                     float* __restrict__ dlnw, float* __restrict__ dlnb) {
    const int col = blockIdx.x * 32 + (threadIdx.x & 31);
    const int rw = threadIdx.x >> 5, nrw = blockDim.x >> 5;
    const int rows_per = (M + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
    float gw = 0.f, gb = 0.f;
    if (col < W) {
        const float wc = w[col], bc = b[col];
        for (int m = r0 + rw; m < r1; m += nrw) {
            const float yh = (y[(long long)m * W + col] - mean[m]) * rstd[m];
            const float z = yh * wc + bc;
            const float sg = sigmoidf_(z);
            const float dz = dx[(long long)m * W + col] * (sg * (1.f + z * (1.f - sg)));
            gw += dz * yh; gb += dz;
        }
    }
    __shared__ float sw[8][33], sb[8][33];
    sw[rw][threadIdx.x & 31] = gw; sb[rw][threadIdx.x & 31] = gb;
    __syncthreads();
    if (rw == 0 && col < W) {
        for (int k = 1; k < nrw; ++k) { gw += sw[k][threadIdx.x]; gb += sb[k][threadIdx.x]; }
        atomicAdd(dlnw + col, gw); atomicAdd(dlnb + col, gb);
    }
}

// out[col] += sum_rows x[row, col]
__global__ void colsum_kernel(int M, int W, const float* __restrict__ x, long long ldx, float* __restrict__ out) {
    const int col = blockIdx.x * 32 + (threadIdx.x & 31);
    const int rw = threadIdx.x >> 5, nrw = blockDim.x >> 5;
    const int rows_per = (M + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
    float s = 0.f;
    if (col < W) for (int m = r0 + rw; m < r1; m += nrw) s += x[(long long)m * ldx + col];
    __shared__ float sh[8][33];
    sh[rw][threadIdx.x & 31] = s;
    __syncthreads();
    if (rw == 0 && col < W) {
        for (int k = 1; k < nrw; ++k) s += sh[k][threadIdx.x];
        atomicAdd(out + col, s);
    }
}

inline int nblk(long long n, int per) { return (int)((n + per - 1) / per); }
inline long long al(long long bytes) { return (bytes + 255) / 256 * 256; }

constexpr int CHUNK_ROWS = 16384;

struct LearnPlan {
    long long total;
    long long r_m, v_m, gae_mask, learn_mask, row_pl, row_ent, row_vl, stats;
    long long Y[D4_MAX_MLP_LAYERS], X[D4_MAX_MLP_LAYERS], mean[D4_MAX_MLP_LAYERS], rstd[D4_MAX_MLP_LAYERS];
    long long out, logits, dlogits, g0, g1;
    long long dyT, xT_hi, xT_lo;      // tensor-core backward: transposed dy (H, Rc) and transposed + tf32-split layer input (H, Rc)
    int Rc;
};

LearnPlan plan_learn(const d4_ctx* c, int B, int T) {
    LearnPlan p; memset(&p, 0, sizeof(p));
    const long long R = (long long)B * T;
    p.Rc = (int)std::min<long long>(R, CHUNK_ROWS);
    long long off = 0;
    auto take = [&](long long bytes) { long long o = off; off += al(bytes); return o; };
    p.r_m = take(R * 4); p.v_m = take(R * 4); p.gae_mask = take(R); p.learn_mask = take(R);
    p.row_pl = take(R * 4); p.row_ent = take(R * 4); p.row_vl = take(R * 4); p.stats = take(64);
    const int H = std::max(std::max(c->cfg.policy_hidden, c->cfg.value_hidden), std::max(c->cfg.value_bins, c->D));
    const int nl = std::max(c->cfg.policy_layers, c->cfg.value_layers);
    for (int l = 0; l < nl - 1; ++l) {
        p.Y[l] = take((long long)p.Rc * H * 4); p.X[l + 1] = take((long long)p.Rc * H * 4);
        p.mean[l] = take((long long)p.Rc * 4); p.rstd[l] = take((long long)p.Rc * 4);
    }
    p.out = take((long long)p.Rc * H * 4);
    const int ldl = std::max(c->ldlog, (c->cfg.value_bins + 3) / 4 * 4);
    p.logits = take((long long)p.Rc * ldl * 4); p.dlogits = take((long long)p.Rc * ldl * 4);
    p.g0 = take((long long)p.Rc * H * 4); p.g1 = take((long long)p.Rc * H * 4);
    if (d4_prec_split(c->cfg.precision)) {
        p.dyT = take((long long)p.Rc * H * 4); p.xT_hi = take((long long)p.Rc * H * 4); p.xT_lo = take((long long)p.Rc * H * 4);
    }
    p.total = off;
    return p;
}

struct MlpGrads { float* const* w; float* const* b; float* const* lnw; float* const* lnb; };

// out[c][r] = in[r][c] for an (R, C) row-major matrix with leading dim ld; out rows have leading dim R.
// SPLIT: also emits the tf32 hi / lo words (out = hi, out_lo = in - hi) so the transposed matrix can be the pre-split
// "weight" operand of the 3xTF32 tensor-core GEMM.
template <bool SPLIT>
__global__ void transpose_kernel(int R, int C, const float* __restrict__ in, long long ld, float* __restrict__ out, float* __restrict__ out_lo) {
    __shared__ float tile[32][33];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;        // 32 x 8 threads
#pragma unroll
    for (int i = ty; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + tx;
        tile[i][tx] = (r < R && c < C) ? in[(long long)r * ld + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + tx;
        if (c < C && r < R) {
            const float v = tile[tx][i];
            if (SPLIT) {
                const float hi = tf32_rna(v);
                out[(long long)c * R + r] = hi; out_lo[(long long)c * R + r] = tf32_rna(v - hi);
            } else {
                out[(long long)c * R + r] = v;
            }
        }
    }
}
int transpose_rows(int R, int C, const float* in, long long ld, float* out, float* out_lo, cudaStream_t s) {
    dim3 grid((C + 31) / 32, (R + 31) / 32);
    if (out_lo) transpose_kernel<true><<<grid, 256, 0, s>>>(R, C, in, ld, out, out_lo);
    else        transpose_kernel<false><<<grid, 256, 0, s>>>(R, C, in, ld, out, nullptr);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError());
    return 0;
}
inline bool learn_tc(const d4_ctx* c) { return d4_prec_split(c->cfg.precision); }   // f16x3 mode: the heads stay on 3xTF32

// forward of one head on a chunk of rows, keeping what the backward needs
int mlp_forward_saved(d4_ctx* c, const MlpW& mlp, const float* x0, int Rc, unsigned char* ws, const LearnPlan& p, float* out, long long ldo,
                      cudaStream_t s) {
    const float* cur = x0; long long ldc = mlp.dims[0];
    for (int l = 0; l < mlp.layers; ++l) {
        const bool last = (l == mlp.layers - 1);
        const int wout = mlp.dims[l + 1];
        float* dst = last ? out : reinterpret_cast<float*>(ws + p.Y[l]);
        const long long ldd = last ? ldo : wout;
        GemmArgs g = gemm_args(cur, ldc, mlp.w[l], mlp.dims[l], dst, ldd, Rc, wout, mlp.dims[l]);
        g.bias = mlp.b[l];
        if (learn_tc(c) && mlp.hi[l] && mlp.lo[l]) {          // 3xTF32 on the tensor cores (fp32-accurate), else exact-fp32 FMA
            LinW lw; lw.w = mlp.w[l]; lw.hi = mlp.hi[l]; lw.lo = mlp.lo[l];
            D4_TRY(d4_engine_gemm(c, g, lw, 0, s));
        } else {
            D4_TRY(d4_gemm_simt(g, s));
        }
        if (!last) {
            float* xn = reinterpret_cast<float*>(ws + p.X[l + 1]);
            D4_TRY(d4_ln_act_rows(dst, wout, mlp.lnw[l], mlp.lnb[l], Rc, wout, xn, wout, D4_ACT_SILU,
                                  reinterpret_cast<float*>(ws + p.mean[l]), reinterpret_cast<float*>(ws + p.rstd[l]), s));
            cur = xn; ldc = wout;
        }
    }
    return 0;
}

// backward of one head on a chunk: dout (Rc, dims[L]) with leading dim ldd is consumed (overwritten)
int mlp_backward(d4_ctx* c, const MlpW& mlp, const float* x0, int Rc, unsigned char* ws, const LearnPlan& p, float* dout, long long ldd,
                 const MlpGrads& G, cudaStream_t s) {
    float* dy = dout; long long ldy = ldd;
    for (int l = mlp.layers - 1; l >= 0; --l) {
        const int wout = mlp.dims[l + 1], win = mlp.dims[l];
        if (l < mlp.layers - 1) {
            const float* Y = reinterpret_cast<float*>(ws + p.Y[l]);
            const float* mean = reinterpret_cast<float*>(ws + p.mean[l]); const float* rstd = reinterpret_cast<float*>(ws + p.rstd[l]);
            dim3 grid(nblk(wout, 32), std::min(64, std::max(1, Rc / 64)));
            ln_param_grad_kernel<<<grid, 256, 0, s>>>(Rc, wout, Y, mean, rstd, mlp.lnw[l], mlp.lnb[l], dy, G.lnw[l], G.lnb[l]);
            D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError());
            ln_silu_bwd_rows_kernel<<<nblk(Rc, RPB), 32 * RPB, 0, s>>>(Rc, wout, Y, mean, rstd, mlp.lnw[l], mlp.lnb[l], dy);
            D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError());
        }
        const float* xin = (l == 0) ? x0 : reinterpret_cast<float*>(ws + p.X[l]);
        {   // dW += dy^T @ x_in
            bool done = false;
            if (learn_tc(c) && p.dyT && (Rc % 4) == 0 && Rc >= 32) {
                // K-major operands for tcgen05: dy^T (wout, Rc) and x_in^T (win, Rc) pre-split into tf32 hi / lo
                float* dyT = reinterpret_cast<float*>(ws + p.dyT);
                float* xT_hi = reinterpret_cast<float*>(ws + p.xT_hi); float* xT_lo = reinterpret_cast<float*>(ws + p.xT_lo);
                GemmArgs g = gemm_args(dyT, Rc, xT_hi, Rc, G.w[l], win, wout, win, Rc);
                g.residual = G.w[l]; g.ldr = win;
                if (d4_gemm_tc_supported(g)) {
                    D4_TRY(transpose_rows(Rc, wout, dy, ldy, dyT, nullptr, s));
                    D4_TRY(transpose_rows(Rc, win, xin, win, xT_hi, xT_lo, s));
                    LinW lw; lw.w = xT_hi; lw.hi = xT_hi; lw.lo = xT_lo;
                    D4_TRY(d4_engine_gemm(c, g, lw, 0, s));
                    done = true;
                }
            }
            if (!done) {
                GemmArgs g = gemm_args(dy, ldy, xin, win, G.w[l], win, wout, win, Rc);
                g.transA = 1; g.transW = 1; g.residual = G.w[l]; g.ldr = win;
                D4_TRY(d4_gemm_simt(g, s));
            }
        }
        {   // db += colsum(dy)
            dim3 grid(nblk(wout, 32), std::min(64, std::max(1, Rc / 64)));
            colsum_kernel<<<grid, 256, 0, s>>>(Rc, wout, dy, ldy, G.b[l]);
            D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError());
        }
        if (l > 0) {   // dx_in = dy @ W
            float* dxin = (dy == reinterpret_cast<float*>(ws + p.g0)) ? reinterpret_cast<float*>(ws + p.g1) : reinterpret_cast<float*>(ws + p.g0);
            bool done = false;
            if (learn_tc(c) && mlp.wthi[l] && mlp.wtlo[l]) {      // dy @ W == dy @ (W^T)^T with W^T (win, wout) K-major
                GemmArgs g = gemm_args(dy, ldy, mlp.wthi[l], wout, dxin, win, Rc, win, wout);
                if (d4_gemm_tc_supported(g)) {
                    LinW lw; lw.w = mlp.wthi[l]; lw.hi = mlp.wthi[l]; lw.lo = mlp.wtlo[l];
                    D4_TRY(d4_engine_gemm(c, g, lw, 0, s));
                    done = true;
                }
            }
            if (!done) {
                GemmArgs g = gemm_args(dy, ldy, mlp.w[l], win, dxin, win, Rc, win, wout);
                g.transW = 1;
                D4_TRY(d4_gemm_simt(g, s));
            }
            dy = dxin; ldy = win;
        }
    }
    return 0;
}

}  // namespace

extern "C" int d4_gae(int B, int T, const float* rewards, const float* values, const uint8_t* masks, const uint8_t* learn_masks,
                      float gamma, float lam, float* returns, void* stream) {
    if (B <= 0 || T <= 0) return 0;
    gae_kernel<<<nblk(B, RPB), 32 * RPB, 0, static_cast<cudaStream_t>(stream)>>>(B, T, rewards, values, masks, learn_masks, gamma, lam, returns);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int64_t d4_learn_workspace_bytes(const d4_ctx* c, int B, int T) {
    if (!c || B <= 0 || T <= 0) return 0;
    return plan_learn(c, B, T).total;
}

extern "C" int d4_learn(d4_ctx* c, const d4_learn_io* io, void* workspace, int64_t workspace_bytes, void* stream) {
    if (!c || !io) return d4_fail("d4_learn: null argument");
    if (!c->bound) return d4_fail("d4_learn: weights not bound");
    if (!c->has_actions) return d4_fail("d4_learn: the model has no discrete actions");
    const int B = io->B, T = io->T;
    if (B <= 0 || T <= 0) return d4_fail("d4_learn: empty experience");
    if (io->objective < D4_OBJECTIVE_PPO || io->objective > D4_OBJECTIVE_PMPO) return d4_fail("d4_learn: unknown objective %d", io->objective);
    if (io->objective == D4_OBJECTIVE_PMPO && io->pmpo_kl_div_loss_weight > 0.f &&
        (!io->old_action_unembeds || io->old_action_unembeds_ld < c->A_total))
        return d4_fail("d4_learn: the pmpo KL term needs old_action_unembeds (B, T, >=%d)", c->A_total);
    const LearnPlan p = plan_learn(c, B, T);
    if (workspace_bytes < p.total) return d4_fail("d4_learn: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)p.total);
    if (reinterpret_cast<uintptr_t>(workspace) & 255) return d4_fail("d4_learn: workspace must be 256-byte aligned");
    if (!c->ws) return d4_fail("d4_learn: d4_set_buffers() must have been called (action table lives in the workspace)");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    const long long R = (long long)B * T;
    float* r_m = reinterpret_cast<float*>(ws + p.r_m); float* v_m = reinterpret_cast<float*>(ws + p.v_m);
    unsigned char* gmask = ws + p.gae_mask; unsigned char* lmask = ws + p.learn_mask;
    float* stats = reinterpret_cast<float*>(ws + p.stats);
    float* row_pl = reinterpret_cast<float*>(ws + p.row_pl); float* row_ent = reinterpret_cast<float*>(ws + p.row_ent);
    float* row_vl = reinterpret_cast<float*>(ws + p.row_vl);

    learn_masks_kernel<<<nblk(R, 256), 256, 0, s>>>(B, T, io->rewards, io->old_values, reinterpret_cast<const long long*>(io->lens),
                                                    io->is_truncated, r_m, v_m, gmask, lmask);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError());
    gae_kernel<<<nblk(B, RPB), 32 * RPB, 0, s>>>(B, T, r_m, v_m, gmask, lmask, io->gamma, io->lam, io->returns);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError());
    adv_stats_kernel<<<1, 1024, 0, s>>>(R, io->returns, v_m, lmask, io->normalize_advantages, io->zscore_eps, io->returns_ema, io->advantages, stats);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError());

    const int D = c->D;
    const int ldl = std::max(c->ldlog, (c->cfg.value_bins + 3) / 4 * 4);
    float* out = reinterpret_cast<float*>(ws + p.out);
    float* logits = reinterpret_cast<float*>(ws + p.logits); float* dlogits = reinterpret_cast<float*>(ws + p.dlogits);
    float* g0 = reinterpret_cast<float*>(ws + p.g0);
    const MlpGrads GP = {io->grad_policy_w, io->grad_policy_b, io->grad_policy_lnw, io->grad_policy_lnb};
    const MlpGrads GV = {io->grad_value_w, io->grad_value_b, io->grad_value_lnw, io->grad_value_lnb};
    const int PH = c->cfg.policy_hidden;

    for (long long r0 = 0; r0 < R; r0 += p.Rc) {
        const int Rc = (int)std::min<long long>(p.Rc, R - r0);
        const float* x0 = io->agent_embed + r0 * D;
        // ---------------- policy head
        D4_TRY(mlp_forward_saved(c, c->policy, x0, Rc, ws, p, out, PH, s));
        {
            GemmArgs g = gemm_args(out, PH, c->unembed, c->unembed_ld, logits, ldl, Rc, c->A_total, PH);
            D4_TRY(d4_gemm_simt(g, s));
        }
        PpoArgs a; memset(&a, 0, sizeof(a));
        a.R = Rc; a.na = c->na; a.A_total = c->A_total; a.sizes_offs = c->b.sizes_offs; a.logits = logits; a.ld = ldl;
        a.actions = reinterpret_cast<const long long*>(io->actions) + r0 * c->na; a.old_logp = io->old_log_probs + r0 * c->na;
        a.adv = io->advantages + r0; a.mask = lmask + r0; a.stats = stats;
        a.eps_clip = io->eps_clip; a.entropy_weight = io->entropy_weight; a.delight_temp = io->delight_temperature; a.use_gate = io->use_delight_gating;
        a.objective = io->objective; a.pmpo_alpha = io->pmpo_pos_to_neg_weight; a.pmpo_kl_weight = io->pmpo_kl_div_loss_weight;
        a.pmpo_reverse_kl = io->pmpo_reverse_kl;
        a.old_logits = io->old_action_unembeds ? io->old_action_unembeds + r0 * io->old_action_unembeds_ld : nullptr;
        a.ld_old = io->old_action_unembeds_ld;
        a.dlogits = dlogits; a.row_pl = row_pl + r0; a.row_ent = row_ent + r0;
        ppo_row_kernel<<<nblk(Rc, RPB), 32 * RPB, 0, s>>>(a);
        D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError());
        {   // d(unembed) += dlogits^T @ policy_embed ; d(policy_embed) = dlogits @ unembed
            GemmArgs g = gemm_args(dlogits, ldl, out, PH, io->grad_unembed, io->grad_unembed_ld, c->A_total, PH, Rc);
            g.transA = 1; g.transW = 1; g.residual = io->grad_unembed; g.ldr = io->grad_unembed_ld;
            D4_TRY(d4_gemm_simt(g, s));
            GemmArgs g2 = gemm_args(dlogits, ldl, c->unembed, c->unembed_ld, g0, PH, Rc, PH, c->A_total);
            g2.transW = 1;
            D4_TRY(d4_gemm_simt(g2, s));
        }
        D4_TRY(mlp_backward(c, c->policy, x0, Rc, ws, p, g0, PH, GP, s));
        // ---------------- value head
        const int K = c->cfg.value_bins;
        D4_TRY(mlp_forward_saved(c, c->value, x0, Rc, ws, p, logits, ldl, s));
        ValueArgs v; memset(&v, 0, sizeof(v));
        v.R = Rc; v.K = K; v.bins = logits; v.ld = ldl; v.returns = io->returns + r0; v.mask = lmask + r0; v.stats = stats;
        v.support = io->value_support; v.sigma_sqrt2 = io->value_sigma_sqrt2; v.hl_eps = io->hl_eps; v.lo = io->value_lo; v.hi = io->value_hi;
        v.dbins = dlogits; v.row_vl = row_vl + r0;
        value_row_kernel<<<nblk(Rc, RPB), 32 * RPB, 0, s>>>(v);
        D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError());
        D4_TRY(mlp_backward(c, c->value, x0, Rc, ws, p, dlogits, ldl, GV, s));
    }
    loss_reduce_kernel<<<1, 1024, 0, s>>>(R, row_pl, row_ent, row_vl, stats, io->entropy_weight, io->losses);
    D4_COUNT_LAUNCH(); D4_CUDA_OK(cudaGetLastError());
    return 0;
}
