/* d4b200.h — C-ABI of the B200-native Dreamer-4 imagination hot path.
 *
 * The reference (lucidrains/dreamer4) is pure Python/PyTorch and has no FFI/plugin boundary of its own
 * (SURVEY.md section 8b); this header is the boundary a maintainer would bind from Python with ctypes.
 * Each entry point names the reference code it replaces (paths relative to the reference repository,
 * D4 = dreamer4/dreamer4.py).  INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch tensors): weights, KV cache, workspace,
 *     inputs and outputs.  The library never allocates device memory and never synchronises the stream.
 *   - every call takes the cudaStream_t to enqueue on (as void*), returns 0 on success or a negative code;
 *     d4_last_error() returns a thread-local message.  Nothing throws across the boundary.
 *   - tensors are fp32 contiguous unless a leading dimension / stride argument says otherwise; action indices and
 *     lens are int64 (torch.long), terminal flags are uint8 (torch.bool storage).
 *   - a d4_ctx is not thread-safe; one ctx per (model, device).
 */
#ifndef D4B200_H
#define D4B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define D4_MAX_ACTION_TYPES 8
#define D4_MAX_MLP_LAYERS 8

/* precision of the dense layers: exact fp32 FMA, tcgen05 TF32, or tcgen05 3xTF32 split (fp32-accurate) */
enum { D4_PREC_FP32 = 0, D4_PREC_TF32 = 1, D4_PREC_TF32X3 = 2,
       D4_PREC_F16X3 = 3 /* experimental (never run on hardware in round 1): 3-term fp16 split on kind::f16 for the transformer's
                            dense layers, 3xTF32 for the heads and d4_learn; see dreamer4_b200/csrc/gemm_f16.cu */ };

/* Mirrors the subset of DynamicsWorldModel.__init__ kwargs (D4:4662-4778) that shapes the imagination path. */
typedef struct d4_config {
    int32_t dim, dim_latent, num_latent_tokens, num_spatial_tokens, num_register_tokens;
    int32_t depth, time_block_every;
    int32_t heads, query_heads, dim_head;            /* attn_heads, attn_kwargs.query_heads, attn_dim_head */
    int32_t pool_heads, pool_dim_head;               /* AttentionPool: 4 x 64 (D4:2147-2148) */
    int32_t ff_inner, ff_inner_pad, ff_act;          /* int(dim*4*2/3); padded to a multiple of 32; 0 silu 1 gelu */
    int32_t max_steps;                               /* K_max = 64 */
    int32_t num_action_types, action_sizes[D4_MAX_ACTION_TYPES];
    int32_t policy_layers, policy_hidden;            /* Linear count of policy_head (depth + 2), hidden width 4*dim */
    int32_t value_layers, value_hidden;
    int32_t terminal_layers, terminal_hidden, predict_terminals;
    int32_t reward_bins, value_bins;
    int32_t num_tasks;
    float   softclamp;                               /* attn_softclamp_value (50) */
    int32_t max_batch, max_time;                     /* KV cache / workspace capacity */
    int32_t precision;                               /* D4_PREC_* of every dense layer (transformer, heads, d4_learn); the final
                                                        action unembedding and ragged / tiny GEMMs always run exact fp32 FMA */
    int32_t time_attn_variant;                       /* K1: 1 cp.async.bulk ring (default of the host class), 0 ld.global staged */
} d4_config;

typedef struct d4_ctx d4_ctx;

const char* d4_last_error(void);
int d4_version(void);
/* number of CUDA kernels this library has launched in the calling process so far */
int64_t d4_launch_count(void);

/* ---- lifetime.  Replaces module construction state that the pass needs at run time (D4:4779-5269). */
int d4_ctx_create(const d4_config* cfg, d4_ctx** out);
void d4_ctx_destroy(d4_ctx* ctx);

/* Packed weights are registered by name (borrowed pointers; see dreamer4_b200/packing.py for the name list and
 * how each is derived from the reference state_dict).  d4_bind() resolves every name the configuration needs
 * and fails with the first missing one. */
int d4_set_weight(d4_ctx* ctx, const char* name, const float* dev_ptr, int64_t numel);
/* D4_PREC_F16X3 only (experimental): a GEMM weight `name` may also be registered as `name.h16hi` / `name.h16lo` - fp16 arrays of
 * the same (out, in) shape holding hi = fp16(q W), lo = fp16(q W - hi), q a power of two that brings rms(q W) to ~1 (dev_ptr
 * cast to const float*, numel counted in ELEMENTS) - together with scale = 1 / q here.  Weights without them, and shapes the
 * fp16 kernel does not take, stay on the 3xTF32 kernel (their `.hi` / `.lo` are required in this mode as well). */
int d4_set_weight_scale(d4_ctx* ctx, const char* name, float scale);
int d4_bind(d4_ctx* ctx);

/* Workspace (activations of one pass) and the time-KV cache are caller-allocated.
 * KV layout: (time_layers, 2, max_batch * S, heads, max_time, dim_head) fp32 — the reference's
 * next_kv_cache layout (D4:3255-3265) with the time axis preallocated; a [:, :, :, :, :T] view of it is a
 * valid reference time_cache. */
int64_t d4_workspace_bytes(const d4_ctx* ctx);
int64_t d4_kv_bytes(const d4_ctx* ctx);
int d4_set_buffers(d4_ctx* ctx, void* workspace, int64_t workspace_bytes, float* kv, int64_t kv_bytes);

/* ---- the hot path */

/* One transformer pass for the newest frame over the KV cache: replaces DynamicsWorldModel.forward's inference
 * branch + get_prediction (D4:6792-7295) and AxialSpaceTimeTransformer.forward (D4:2927-3267) for time == 1.
 *   latent        (B, N, Dl)   the (noised) latent of the new frame
 *   prev_actions  (B, na) int64 row stride pa_stride, or NULL at frame 0 (D4:7105-7126)
 *   tasks         (B) int64 or NULL
 *   t             number of frames already in the cache (= rotary offset, D4:3010)
 *   commit_kv     1 on the clean pass: append this frame's keys/values at position t (D4:6545-6546)
 *   pred_out      (B, N, Dl)   predicted clean latent (D4:7251)
 *   agent_out     (B, D)       agent-token embedding (D4:7222, 7279) */
int d4_pass(d4_ctx* ctx, int B, const float* latent, int signal_level, int step_size_log2,
            const int64_t* prev_actions, int64_t pa_stride, const int64_t* tasks, int t, int commit_kv,
            float* pred_out, float* agent_out, void* stream);

typedef struct d4_frame_io {
    /* inputs */
    const float* noise_latent;      /* (B, N, Dl)  randn, D4:6475 */
    const float* action_uniform;    /* (B, A_total) rand for gumbel-argmax, D4:6637 / MultiCategorical.sample; NULL if no actions */
    const float* terminal_uniform;  /* (B) rand for the Bernoulli terminal draw, D4:6611; NULL unless terminals */
    const int64_t* prev_actions; int64_t pa_stride;   /* previous frame's actions (B, na), NULL at frame 0 */
    const int64_t* tasks;
    /* outputs: rows of the (B, T, ...) Experience tensors for this frame, each with its batch stride (elements) */
    float* latents;      int64_t latents_bs;     /* (N*Dl) per b, clamped to [-1, 1] (D4:6686) */
    float* agent_embed;  int64_t agent_bs;       /* (D) */
    float* rewards;      int64_t rewards_bs;     /* scalar */
    float* values;       int64_t values_bs;      /* scalar */
    int64_t* actions;    int64_t actions_bs;     /* (na) */
    float* log_probs;    int64_t log_probs_bs;   /* (na) */
    float* logits;       int64_t logits_bs;      /* (A_total)  = old_action_unembeds (D4:6749-6750) */
    int64_t* lens;  uint8_t* terminals;          /* (B) updated in place when terminals are predicted */
} d4_frame_io;

/* One imagined frame: the num_steps denoising passes + the clean pass + reward / terminal / policy / value heads
 * + action sampling.  Replaces one iteration of the frame loop of DynamicsWorldModel.generate (D4:6458-6684). */
int d4_frame(d4_ctx* ctx, int B, int t, int num_steps, float discrete_temperature, const d4_frame_io* io, void* stream);

/* One OBSERVED frame: io->noise_latent holds the clean latent of a real observation (not noise); runs only the clean pass at
 * signal level max_steps-1 with the step-size embedding of `num_steps` (appending the frame's keys/values at position t) and
 * the same heads as d4_frame.  Replaces one iteration of DynamicsWorldModel.interact_with_env's loop between the tokenizer
 * and env.step (D4:5612-5673).  io->latents receives the clamped copy of the input (scratch for most callers). */
int d4_observe(d4_ctx* ctx, int B, int t, int num_steps, float discrete_temperature, const d4_frame_io* io, void* stream);

/* ---- in-situ kernel timing (bench.py's roofline): CUDA events recorded around every launch of the four kernel
 * classes on the launching stream while enabled.  d4_profile_read() synchronises the recorded events and returns,
 * per class [gemm, time_attn (K1), small_attn, other-marked], {milliseconds, launches, algorithmic work} where work is
 * FLOPs for gemm (2*M*N*K) and bytes for time_attn (M*h*d*4*(2t+4), +2 rows on the append pass), 0 otherwise. */
#define D4_PROF_CLASSES 4
int d4_profile(d4_ctx* ctx, int enable);
int d4_profile_read(d4_ctx* ctx, double* out /* [D4_PROF_CLASSES][3] */);

/* ---- stand-alone operators (unit parity tests, and the reference's attend_fn / calc_gae seams) */

/* K1 stand-alone: naive_attend for a single new query per (token, head) over cache + self (D4:1683-1756, 2021-2054).
 * qkvgm rows: [q (hq*d) | k (h*d) | v (h*d) | gate logits (hq) | mix logits (h)], leading dim ld. */
int d4_time_attn_decode(int M, int heads, int query_heads, int dim_head, int t, int Tmax,
                        const float* qkvgm, int64_t ld, const float* v0, const float* k_gamma, const float* inv_freq,
                        float* kcache, float* vcache, float* out, float softclamp, int commit, int variant, void* stream);

/* C = epilogue(A @ W^T): precision D4_PREC_*; act 0 none / 1 GLU-silu / 2 GLU-gelu (W rows interleaved x,g).
 * D4_PREC_F16X3: W / W_lo point to fp16 (N, ldw) arrays hi = fp16(q W), lo = fp16(q W - hi) with q a power of two that brings
 * rms(q W) to ~1; the caller folds 1 / q into row_scale. */
int d4_linear(int precision, int M, int N, int K, const float* A, int64_t lda, const float* W, int64_t ldw, const float* W_lo,
              const float* bias, const float* row_scale, const float* residual, int64_t ldr, int act,
              float* C, int64_t ldc, void* stream);

/* d4_pass with per-dream conditioning: signal_levels (B) int64 in [0, max_steps), step_sizes_log2 (B) int64 - the (b, t) / (b)
 * tensors DynamicsWorldModel.forward takes (D4:6792-6827, 6912-6942), one frame t of them per call.  Replaces one frame of the
 * inference branch of DynamicsWorldModel.forward + get_prediction (D4:6792-7295) for latent_is_noised = True. */
int d4_pass_ex(d4_ctx* ctx, int B, const float* latent, const int64_t* signal_levels, const int64_t* step_sizes_log2,
               const int64_t* prev_actions, int64_t pa_stride, const int64_t* tasks, int t, int commit_kv, float* pred_out,
               float* agent_out, void* stream);

/* A head MLP on caller rows: which = 0 policy_head (D -> 4D), 1 value_head (D -> value bins), 2 terminal head (Dl -> 1); x (M, dim_in)
 * contiguous, out (M, dim_out) contiguous, M <= max_batch.  Replaces calling the reference's head modules directly
 * (e.g. dynamics.policy_head(embeds.agent), tests/test_dreamer.py:1262, x-mlps create_mlp D4:4950-4956). */
int d4_head_forward(d4_ctx* ctx, int which, const float* x, int M, float* out, void* stream);

/* ---- diagnostics (bench / profiling scripts and tests only).
 * d4_graph_replays: how many frames of this context ran as a CUDA-graph replay (0 = every frame was launched directly).
 * d4_debug_set: switches of individual kernels: ("gemm_f16", bits) - timing ablations of gemm_f16.cu, results are garbage for bits
 *   1 | 2 | 4 | 8 | 16 | 128 (64 = residual chunk loads issued late: exact); ("lp_fused", 1 | 2) - per-frame / persistent space -> latent
 *   pool kernel (bit-identical, tests/test_horizon_parity_gpu.py). */
int64_t d4_graph_replays(const d4_ctx* ctx);
int d4_debug_set(const char* key, int value);
/* counters of a context: "graph_enabled", "graph_keys", "graph_captured", "graph_capture_refused"; -1 for an unknown key */
int64_t d4_debug_get(const d4_ctx* ctx, const char* key);

/* calc_gae (D4:1566-1600): returns = reverse-scan(delta, gamma*lambda*mask) + values.  masks/learn_masks uint8 (B,T). */
int d4_gae(int B, int T, const float* rewards, const float* values, const uint8_t* masks, const uint8_t* learn_masks,
           float gamma, float lam, float* returns, void* stream);

/* ---- learn_from_experience (D4:5893-6305), objective 'ppo' | 'spo' | 'pmpo' (D4:6127-6212), heads only.
 * Computes both losses and the gradients of every policy-head / unembed / value-head parameter in one call. */
#define D4_OBJECTIVE_PPO  0
#define D4_OBJECTIVE_SPO  1
#define D4_OBJECTIVE_PMPO 2
typedef struct d4_learn_io {
    int32_t B, T;
    const float* agent_embed;       /* (B, T, D) */
    const float* rewards;           /* (B, T) */
    const float* old_values;        /* (B, T) */
    const int64_t* actions;         /* (B, T, na) */
    const float* old_log_probs;     /* (B, T, na) */
    const int64_t* lens;            /* (B) */
    const uint8_t* is_truncated;    /* (B) */
    const uint8_t* terminals;       /* (B) or NULL */
    float gamma, lam, eps_clip, entropy_weight, delight_temperature, zscore_eps;
    int32_t use_delight_gating, normalize_advantages;
    const float* value_support;     /* (value_bins + 1) HL-Gauss bin edges */
    float value_sigma_sqrt2, hl_eps, value_lo, value_hi;
    /* outputs */
    float* losses;                  /* [policy_total, value, policy_surrogate, entropy_term] */
    float* returns;                 /* (B, T) */
    float* advantages;              /* (B, T) normalised */
    /* gradients, in the order policy layers (W, b, ln_w, ln_b)..., unembed, value layers...; see packing.py */
    float* grad_policy_w[D4_MAX_MLP_LAYERS]; float* grad_policy_b[D4_MAX_MLP_LAYERS];
    float* grad_policy_lnw[D4_MAX_MLP_LAYERS]; float* grad_policy_lnb[D4_MAX_MLP_LAYERS];
    float* grad_unembed; int64_t grad_unembed_ld;   /* (A_total, 4D) rows with leading dim (the [:, 0] slice of (A, mtp, 4D)) */
    float* grad_value_w[D4_MAX_MLP_LAYERS]; float* grad_value_b[D4_MAX_MLP_LAYERS];
    float* grad_value_lnw[D4_MAX_MLP_LAYERS]; float* grad_value_lnb[D4_MAX_MLP_LAYERS];
    /* surrogate objective (zero-initialised = ppo).  pmpo (D4:6127-6182): weight of the sign-split log-likelihood term,
     * weight and direction of the KL to the logits stored at rollout time (Experience.old_action_unembeds, D4:5868-5869);
     * the KL is one categorical over the flat A_total logits, as the reference calls it (D4:6160-6169). */
    int32_t objective;              /* D4_OBJECTIVE_* */
    int32_t pmpo_reverse_kl;        /* 1: KL(old || new) (reference default), 0: KL(new || old) */
    float pmpo_pos_to_neg_weight, pmpo_kl_div_loss_weight;
    const float* old_action_unembeds; int64_t old_action_unembeds_ld;   /* (B, T, A_total) rows with leading dim; pmpo only */
    /* keep_reward_ema_stats (D4:5987-6013): device pointer to [ema_returns_mean, ema_returns_std] (already updated with this
     * batch's returns by the caller) or NULL; advantage = (returns - mean) / std - (old_values - mean) / std. */
    const float* returns_ema;
} d4_learn_io;

int64_t d4_learn_workspace_bytes(const d4_ctx* ctx, int B, int T);
int d4_learn(d4_ctx* ctx, const d4_learn_io* io, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- video tokenizer (VideoTokenizer.tokenize / .decode, D4:4107-4113, 4183-4237; both of its
 * transformers run one frame per step over a time-KV cache, like the dynamics model).  STATUS: drafted in round 1, not yet run
 * on hardware. */

/* A generic AxialSpaceTimeTransformer context: the tokenizer's encoder (D4:3908-3929: num_special = num_latent_tokens) or decoder
 * (D4:3595-3607: library defaults, num_special = 1); S = tokens_per_frame = patches + latent tokens, the special tokens last.
 * The returned d4_ctx is used with d4_set_weight / d4_bind / d4_workspace_bytes / d4_kv_bytes / d4_set_buffers / d4_ctx_destroy
 * exactly like a d4_ctx_create one; packed weight names: vr.w, inv_freq, final_norm, L{i}.attn.*, L{i}.ff.*, P{i}.*, PF.*, FA.*,
 * FAFF.* (dreamer4_b200/packing.py: _pack_transformer). */
typedef struct d4_tf_config {
    int32_t dim, depth, time_block_every;
    int32_t heads, query_heads, dim_head, pool_heads, pool_dim_head;
    int32_t ff_inner, ff_inner_pad, ff_act;
    int32_t tokens_per_frame, num_special, final_norm;
    float   softclamp;
    int32_t max_batch, max_time, precision, time_attn_variant;
} d4_tf_config;
int d4_tf_create(const d4_tf_config* cfg, d4_ctx** out);

/* One frame through the transformer over its time-KV cache (appended at position t): replaces
 * AxialSpaceTimeTransformer.forward (D4:2927-3267) for time == 1.  tokens_in / tokens_out (B, S, D). */
int d4_tf_step(d4_ctx* ctx, int B, const float* tokens_in, int t, float* tokens_out, void* stream);

/* C (M, N) = A[rows] @ W^T + bias, the rows of A taken through a grouped row map: compact row m -> (m / a_grp) * a_gstride +
 * a_goff + m % a_grp (a_grp = 0: identity) - e.g. the latent-token rows of every frame of a (B, S, D) token tensor.
 * W / W_lo: the tf32 hi / lo words (D4_PREC_TF32X3) or W itself; W_exact: the fp32 weight, used by the exact-fp32 kernel when
 * precision is D4_PREC_FP32 or the shape does not fit the tensor-core path.  Replaces the nn.Linear layers either side of the
 * tokenizer's transformers (D4:3836, 3884, 4413, 4148, 3569). */
int d4_linear_rows(int precision, int M, int N, int K, const float* A, int64_t lda, int a_grp, int a_gstride, int a_goff,
                   const float* W, int64_t ldw, const float* W_lo, const float* W_exact, const float* bias, float* C, int64_t ldc,
                   void* stream);

/* 'b c (h p1) (w p2) -> (b h w) (p1 p2 c)' of one frame (D4:3835, 3883); frame element (b, c, y, x) at frame[b * stride_b +
 * c * stride_c + y * W + x] (a [:, :, t] slice of a (b c t h w) video); out (B * H/p * W/p, p * p * C). */
int d4_patchify(int B, int C, int H, int W, int p, const float* frame, int64_t stride_b, int64_t stride_c, float* out, void* stream);
/* frame <- frame + (pred - frame) * scale with pred the un-patched rows '(b h w) (p1 p2 c) -> b c (h p1) (w p2)' (D4:3571) and
 * scale = 1 / (1 - tau) / steps: one Euler step of the decoder's flow (D4:4223-4227). */
int d4_unpatchify_flow(int B, int C, int H, int W, int p, const float* patches, float* frame, int64_t stride_b, int64_t stride_c,
                       float scale, void* stream);
/* tokens (B, S, D) of one frame: rows i < P = LayerNorm(lin[b * P + i]) * ln_w (nn.LayerNorm(bias=False), D4:3837, 3885)
 * + pos_emb[i] (decoder only, D4:3618-3628; NULL otherwise); rows P .. S-1 = the num_special special tokens
 * special[b * special_bstride + j * D] (special_bstride = 0: one learned set for every b, D4:4349). */
int d4_tok_assemble(int B, int S, int P, int D, const float* lin, const float* ln_w, const float* pos_emb, const float* special,
                    int64_t special_bstride, int num_special, float* tokens, void* stream);
/* x <- tanh(x) over n elements (D4:4426). */
int d4_tanh_rows(float* x, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif
