// Internal launch-wrapper declarations shared by the engine and the C-ABI unit entry points.
#pragma once
#include "common.cuh"

// ---- GEMM (gemm_simt.cu / gemm_tc.cu)
int d4_gemm_simt(const GemmArgs& g, cudaStream_t stream);
// M <= 32 rows: exact-fp32 weight-streaming kernel (gemm_skinny.cu), N / 4 CTAs
int d4_gemm_skinny(const GemmArgs& g, cudaStream_t stream);
int d4_gemm_skinny_supported(const GemmArgs& g);
// tcgen05 path: terms = 1 (tf32) or 3 (tf32x3 split: A split in shared memory, W_lo supplied)
int d4_gemm_tc(const GemmArgs& g, int terms, cudaStream_t stream);
int d4_gemm_tc_supported(const GemmArgs& g);
int d4_gemm_pair_default(void);
int d4_gemm_pair_bn(int M, int N);      // N-tile width (128 | 256) the CTA-pair kernels pick for an (M, N) product
// persistent warp-specialised tcgen05 kernel (gemm_tc2.cu); bn = 128 / 256 / 0 (auto)
int d4_gemm_tc2(const GemmArgs& g, int terms, int bn, cudaStream_t stream);
// CTA-pair (cta_group::2) persistent kernel (gemm_tc3.cu): 256 x bn output tiles, bn = 128 / 256 / 0 (auto)
int d4_gemm_tc3(const GemmArgs& g, int terms, int bn, cudaStream_t stream);
// experimental fp16 3-term split (gemm_f16.cu): g.W / g.W_lo are fp16 (N, ldw) arrays of the pre-scaled weights, w_scale = 1 / q
int d4_gemm_f16x3(const GemmArgs& g, float w_scale, int bn, cudaStream_t stream);
int d4_gemm_f16x3_supported(const GemmArgs& g, const void* whi, const void* wlo);

// ---- row-wise kernels (rowops.cu)
struct AssembleArgs {
    float* tokens;               // (B, S, D)
    int B, S, D, nsp, nreg, has_actions, na;
    const float* sig_emb; const float* step_emb; int signal, step;
    const float* registers;      // (nreg, D)
    const float* agent_embed;    // (D)
    const float* action_learned; // (D)
    const float* action_emb;     // (A_total, D)
    const long long* prev_actions; long long pa_stride;   // (B, na) int64 rows with stride, nullptr at frame 0
    int act_off[8];
    const float* task_emb; const long long* tasks;
    const long long* signal_rows; const long long* step_rows;   // optional (B): per-row signal level / log2 step size (override signal / step)
};
int d4_row_rstd(const float* x, long long ldx, RowMap map, int M, int D, float* out, cudaStream_t s);
int d4_row_sumsq(const float* x, long long ldx, int M, int D, float* out, cudaStream_t s);     // out[m] = sum_d x[m][d]^2
// rows map(0..M-1): out_c[m] = rstd (compact), out_f[map(m)] = sum of squares | rstd (full-row index); either may be null
int d4_row_stat_map(const float* x, long long ldx, RowMap map, int M, int D, float* out_c, float* out_f, int full_is_ss, cudaStream_t s);
int d4_rmsnorm_rows(const float* x, long long ldx, RowMap map, const float* w, int M, int D, float* out, long long ldo, cudaStream_t s);
int d4_ln_act_rows(const float* x, long long ldx, const float* w, const float* b, int M, int D, float* out, long long ldo, int act,
                   float* save_mean, float* save_rstd, cudaStream_t s);
int d4_assemble_tokens(const AssembleArgs& a, cudaStream_t s);
int d4_flow_step(float* x, const float* pred, long long n, float one_minus_tau, float dt, cudaStream_t s);
int d4_store_latents(const float* x, float* out, int B, long long per_b, long long out_bstride, cudaStream_t s);
int d4_copy_rows(const float* src, long long lds, float* dst, long long ldd, int M, int D, cudaStream_t s);
int d4_gather_rows(const float* src, long long lds, RowMap map, long long M, int D, float* dst, long long ldd, cudaStream_t s);
int d4_hl_gauss_decode(const float* logits, long long ld, int M, int K, const float* centers, float* out, long long out_stride, cudaStream_t s);
int d4_sample_actions(const float* logits, long long ld, const float* u, long long ldu, int B, int na, const int* sizes_offs, float inv_temp,
                      long long* actions, long long act_stride, float* logp, long long lp_stride, cudaStream_t s);
int d4_mean_tokens(const float* x, int B, int N, int Dl, float* out, cudaStream_t s);
int d4_terminal_update(const float* logit, long long ld, const float* u, int B, int frame, long long* lens, unsigned char* terminals, cudaStream_t s);

// ---- attention (attn.cu)
// Generic small attention: one warp per (batch item, kv head); K/V of the item staged in shared memory.
// Covers the 15-token space attention (softclamp + agent mask + value-residual + belief), the latent<->spatial
// learned-query pools, the attention-residual pools over layer hiddens and the final agent cross-attention.
struct SmallAttnArgs {
    int nb, hkv, g, d, nq, n;
    const float* q; long long q_sb, q_si;          // q(b,i,hq)   = q + b*q_sb + i*q_si + hq*d       (hq = hk*g + gi)
    const float* k; long long k_sb, k_sj;          // k(b,j,hk)   = k + b*k_sb + j*k_sj + hk*d
    const float* v; long long v_sb, v_sj;
    const float* k_gamma;                          // (hkv, d) MultiHeadRMSNorm gamma
    const float* v0; long long v0_sb, v0_sj;       // value residual (addressed like v) or nullptr
    const float* mix; long long mix_sb, mix_sj;    // mix logit(b,j,hk) = mix + b*mix_sb + j*mix_sj + hk
    const float* gate; long long gate_sb, gate_si; // gate logit(b,i,hq) = gate + b*gate_sb + i*gate_si + hq, or nullptr
    float* out; long long out_sb, out_si;          // out(b,i,hq) = out + b*out_sb + i*out_si + hq*d
    float scale, softclamp;
    int mask_agent, belief;
    // attention-residual pools only: when gate_w is set, the head-gate logits are computed in the kernel as
    // gate_rstd[b] * (gate_x[b] . gate_w[h]) instead of being read from `gate` (keeps the 4 gate rows out of the q GEMM,
    // whose N then is exactly 256)
    const float* gate_x; long long gate_x_ld; const float* gate_rstd; const float* gate_w; int gate_D;
    int gate_rstd_is_ss;    // gate_rstd holds the sum of squares of the token row: rstd = rsqrt(ss / gate_D + eps)
    int allow_tensor;       // space attention: 3xTF32 mma.sync tiles allowed (tf32x3 / tf32 engine modes); 0 = exact-fp32 FMA
};
int d4_small_attn(const SmallAttnArgs& a, cudaStream_t s);
// space attention with every operand in registers as MMA fragments (space_attn.cu): head dim 64, S <= 16
int d4_space_attn_reg(const SmallAttnArgs& a, cudaStream_t s);
int d4_space_attn_reg_ok(const SmallAttnArgs& a);
int d4_pool_attn_ok(const SmallAttnArgs& a);      // 1 if `a` takes the one-warp-per-token pool kernel (the only one that honours gate_w)

// K1: time-decode attention over the in-place KV cache (+ append on the clean pass).
struct TimeAttnArgs {
    int M, hkv, g, d, t, Tmax;
    const float* qkvgm; long long ld;              // row: [q (hq*d) | k (hkv*d) | v (hkv*d) | gate logits (hq) | mix logits (hkv)]
    int off_k, off_v, off_g, off_m;
    const float* v0; long long ldv0;               // (M, hkv*d)
    const float* k_gamma;                          // (hkv, d)
    const float* inv_freq;                         // (d/2)
    float* kcache; float* vcache;                  // (M, hkv, Tmax, d) each
    float* out; long long ldo;                     // (M, hq*d)
    float scale, softclamp;
    int commit;
    int variant;                                   // 0 = ld.global staged, 1 = cp.async.bulk ring
    RowMap tmap;                                   // bulk variant: stream m of the M launched covers token tmap(m) (rows of qkvgm / v0 / out and the cache); grp 0 = identity
};
int d4_time_attn(const TimeAttnArgs& a, cudaStream_t s);

// ---- fused learned-query pools (fused_pools.cu)
// latents -> gated attention output of latents_to_spatial_tokens (reference dreamer4.py:2179-2210, 4822-4828)
struct L2sArgs {
    int B, N, Dl, nsp, h, hq, g, d, Dq;
    const float* latent;          // (B, N, Dl)
    const float* w_k;             // (h*d, Dl)  to_k with norm_context gamma folded
    const float* w_v;             // (h*d, Dl)
    const float* q;               // (nsp, Dq)  projected learned queries
    const float* gate;            // (nsp, hq)  gate logits
    const float* k_gamma;         // (h, d)
    float* out;                   // (B, nsp, Dq)
    float scale;
    int allow_tensor;             // key projection on mma.sync 3xTF32 tiles (tensor-core engine modes); 0 = exact-fp32 FMA
};
int d4_l2s_fused_supported(const L2sArgs& a);
int d4_l2s_fused(const L2sArgs& a, cudaStream_t s);
// spatial-token keys/values -> predicted latents of to_latent_pred (reference dreamer4.py:4830-4834)
struct LpArgs {
    int B, N, Dl, nsp, h, hq, d;
    const float* kv;              // (B, nsp, 2*h*d)  [K | V] of the normalised spatial tokens
    const float* q;               // (N, hq*d)  projected learned queries
    const float* gate;            // (N, hq)    gate logits
    const float* k_gamma;         // (h, d)
    const float* w_comb;          // (Dl, hq*d) Linear(D -> Dl) @ to_out
    float* pred;                  // (B, N, Dl)
    float scale;
};
int d4_lp_fused_supported(const LpArgs& a);
int d4_lp_fused(const LpArgs& a, cudaStream_t s);
void d4_lp_fused_debug(int version);

// ---- attention within a frame of more than 64 tokens (frame_attn.cu; video tokenizer).  a.mask_agent = number of special tokens.
int d4_frame_attn(const SmallAttnArgs& a, cudaStream_t s);

// ---- video tokenizer front / back end (tokenizer.cu)
int d4_patchify_launch(int B, int C, int H, int W, int p, const float* frame, long long sb, long long sc, float* out, cudaStream_t s);
int d4_unpatchify_flow_launch(int B, int C, int H, int W, int p, const float* patches, float* frame, long long sb, long long sc, float scale,
                              cudaStream_t s);
int d4_tok_assemble_launch(int B, int S, int P, int D, const float* lin, const float* ln_w, const float* pos_emb, const float* special,
                           long long special_bstride, float* tokens, cudaStream_t s);
int d4_tanh_launch(float* x, long long n, cudaStream_t s);
