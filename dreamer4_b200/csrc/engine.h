// d4_ctx: configuration, bound weight pointers and the workspace plan of one model on one device.
#pragma once
#include <array>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>
#include "../../include/d4b200.h"
#include "kernels.h"

// GEMM weight, its tf32 hi/lo split (tf32x3), and - f16x3 engine mode only - the fp16 hi/lo words of q W with h_scale = 1 / q
struct LinW { const float* w = nullptr; const float* hi = nullptr; const float* lo = nullptr;
              const void* h_hi = nullptr; const void* h_lo = nullptr; float h_scale = 1.f; };
// the two engine modes whose dense layers are fp32-accurate tensor-core split products
static inline bool d4_prec_split(int p) { return p == D4_PREC_TF32X3 || p == D4_PREC_F16X3; }

struct AttnLayerW { LinW w; const float* b; const float* k_gamma; LinW w_out; };
struct FFW { LinW w_in; const float* b_in; LinW w_out; const float* b_out; };
struct PoolW { LinW w_qg; LinW w_kv; const float* k_gamma; LinW w_out; };
struct MlpW { int layers = 0; const float* w[D4_MAX_MLP_LAYERS]; const float* b[D4_MAX_MLP_LAYERS];
              const float* hi[D4_MAX_MLP_LAYERS]; const float* lo[D4_MAX_MLP_LAYERS];   // optional tf32 split of w (tf32x3 heads)
              const void* h_hi[D4_MAX_MLP_LAYERS]; const void* h_lo[D4_MAX_MLP_LAYERS]; float h_scale[D4_MAX_MLP_LAYERS];   // optional fp16 split of q w + 1 / q (f16x3 rollout heads)
              const float* wthi[D4_MAX_MLP_LAYERS]; const float* wtlo[D4_MAX_MLP_LAYERS]; // optional tf32 split of w^T (in, out): learn backward
              const float* lnw[D4_MAX_MLP_LAYERS]; const float* lnb[D4_MAX_MLP_LAYERS]; int dims[D4_MAX_MLP_LAYERS + 1]; };

struct d4_ctx {
    d4_config cfg;
    // derived
    int S, D, Dl, N, nsp, nreg, L, y, h, hq, d, Dq, Dkv, hp, dp, Dp, inner, inner_pad, n_hid;
    int has_actions, na, A_total, same_len;
    int NQ, ldq;            // fused qkv+gates+mix row
    int ldpq;               // pool q+gates row
    int ldfa;               // agent q+gates row
    int ldlog;              // logits row
    std::vector<int> is_time;
    int act_off[D4_MAX_ACTION_TYPES];

    std::unordered_map<std::string, std::pair<const float*, int64_t>> table;
    std::unordered_map<std::string, float> scales;      // d4_set_weight_scale: 1 / q of the fp16-split weights
    bool bound = false;

    // bound weights
    const float *sig_emb, *step_emb, *registers, *agent_embed, *action_learned, *action_emb, *task_emb;
    LinW l2s_w_kv, l2s_w_out, l2s_w; const float *l2s_q, *l2s_gate, *l2s_k_gamma, *l2s_b;
    LinW vr_w;
    std::vector<AttnLayerW> attn; std::vector<FFW> ff; std::vector<PoolW> pools; PoolW pool_final;
    PoolW fa;               // final agent cross-attention (w_qg, w_kv, k_gamma, w_out)
    FFW fa_ff;
    const float *lp_norm0, *lp_norm_ctx, *lp_q, *lp_gate, *lp_k_gamma; LinW lp_w_kv, lp_w_comb, lp_w;
    const float* inv_freq;
    const float *reward_w, *reward_centers, *value_centers;
    MlpW policy, value, terminal;
    const float* unembed; int64_t unembed_ld;

    // buffers
    unsigned char* ws = nullptr; int64_t ws_bytes = 0; float* kv = nullptr; int64_t kv_bytes_ = 0;
    int64_t ws_need = 0;
    struct {
        float *lat_x, *lat_rstd, *kv_l, *att_l, *hid, *hid_rstd, *x_cur, *x_rstd, *qkvgm, *v0, *attn_o, *ff_mid,
              *pool_qg, *pool_kv, *pool_att, *fa_kv, *fa_q, *fa_att, *ag_rstd, *sp_n, *sp_n2, *sp_kv, *lp_att, *pred,
              *hbuf0, *hbuf1, *logits, *bins, *agent, *term_in, *fin_x, *fin_rstd, *fin_ctx_rs, *fin_rstd2;
        int* sizes_offs;
    } b;
    bool sizes_uploaded = false;

    // in-situ profiling (d4_profile / d4_profile_read)
    bool prof_on = false;
    bool fuse_ss = true;         // RMS statistics accumulated in the producing GEMM's epilogue (D4_FUSE_SS=0: separate row pass)
    bool space_mma = true;       // space attention on mma.sync 3xTF32 tiles in the tensor-core engine modes (D4_SPACE_MMA=0: FMA kernel)
    bool trim_final = true;      // final attention-residual pool (and the agent cross-attention) only for the token rows a pass's outputs read (D4_TRIM_FINAL=0: all rows)
    bool trim_cone = true;       // denoise passes: everything after the last space layer's attention on the spatial rows only (D4_TRIM_CONE=0)
    bool skinny = true;          // GEMMs of <= 32 rows on the weight-streaming exact-fp32 kernel (gemm_skinny.cu); D4_SKINNY=0: the tile kernels
    bool fuse_pools = true;      // fused latent<->space pool kernels (fused_pools.cu); D4_FUSE_POOLS=0 keeps the GEMM + attention path
    // generic transformer context (d4_tf_create: the video tokenizer's encoder / decoder): S tokens per frame of which the last
    // tf_ns are special, final RMSNorm; uses attn / ff / pools / pool_final / fa / fa_ff / vr_w / inv_freq and the buffers below
    bool tf_mode = false; int tf_ns = 1; int tf_final_norm = 0; const float* tf_final_norm_w = nullptr;
    struct { float *fa_q, *fa_att, *sp_rstd; } tfb = {};

    // CUDA-graph replay of whole frames (on by default for frames of <= graph_max_rows token rows, D4_GRAPH=0 disables; small
    // batches are launch-bound: ~570 launches per imagined frame).  One instantiated graph per (entry point, B, t, num_steps, temperature, which optional io fields are present),
    // captured the SECOND time a key is seen (the first use runs directly, which also gets every lazy one-time setup out of
    // the way); the graph works on dense staging rows inside the workspace, copied in / out around the launch, so it does not
    // depend on the caller's pointers.  Dropped whenever weights or buffers are re-registered.  Capture and replay run on a stream
    // the engine owns, ordered against the caller's stream by two events, so the caller may use any stream (also the legacy default).
    bool use_graphs = false;
    int graph_max_rows = 4096;   // frames with more than this many token rows (B * S) always run directly
    struct GraphIO { float *noise, *act_u, *term_u, *latents, *agent, *rewards, *values, *logp, *logits;
                     long long *prev_actions, *tasks, *actions, *lens; unsigned char* terminals; } gio = {};
    struct FrameGraph { cudaGraphExec_t exec = nullptr; long long launches = 0; bool seen = false; bool direct = false; };   // direct: capture refused, keep launching
    std::map<std::array<long long, 6>, FrameGraph> graphs;
    long long graph_replays = 0;
    cudaStream_t gstream = nullptr; cudaEvent_t gev_in = nullptr, gev_out = nullptr;    // capture / replay stream, ordered against the caller's by events
    struct ProfRec { cudaEvent_t a, b; int cls; double work; };
    std::vector<ProfRec> prof;          // records in use
    std::vector<ProfRec> prof_pool;     // created events, reused
};

enum { D4_CLS_GEMM = 0, D4_CLS_TIME_ATTN = 1, D4_CLS_SMALL_ATTN = 2, D4_CLS_OTHER = 3 };
// usage: int h = d4_prof_begin(c, cls, work, s); launch...; d4_prof_end(c, h, s);
int d4_prof_begin(d4_ctx* c, int cls, double work, cudaStream_t s);
void d4_prof_end(d4_ctx* c, int handle, cudaStream_t s);

int d4_engine_plan(d4_ctx* c);
int d4_engine_gemm(d4_ctx* c, GemmArgs g, const LinW& w, int force_fp32, cudaStream_t s);
int d4_mlp_forward(d4_ctx* c, const MlpW& mlp, const float* x, long long ldx, int M, float* buf0, float* buf1, float* out, long long ldo,
                   int allow_tensor, cudaStream_t s);
