// The imagination engine: one transformer pass for the newest frame over the in-place time-KV cache, the
// per-frame denoise loop + heads, and the C-ABI entry points declared in include/d4b200.h.
//
// Data layout in HBM (fp32, row-major; M = B*S tokens of the new frame, S = tokens per frame):
//   hid      (2L+1, M, D)   residual-stream snapshots [input, post-attn_0, post-ff_0, ...] — the context of
//                           the attention-residual pools (reference dreamer4.py:3040, 3172, 3216)
//   hid_rstd (2L+1, M)      their RMS statistics (gamma is folded into the consuming GEMM's weight)
//   qkvgm    (M, ldq)       fused projection row [q | k | v | gate logits | value-residual mix logits]
//   kv cache (y, 2, Mmax, h, Tmax, d)  reference next_kv_cache layout (dreamer4.py:3255-3265), time axis preallocated
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <algorithm>
#include "engine.h"

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[1024] = "";
int d4_fail(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
    return -1;
}
int d4_fail_cuda(cudaError_t e, const char* what, const char* file, int line) {
    snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return -2;
}
extern "C" const char* d4_last_error(void) { return g_err; }
extern "C" int d4_version(void) { return 100; }
long long d4_launches_ = 0;
extern "C" int64_t d4_launch_count(void) { return d4_launches_; }
void d4_gemm_f16_debug(int bits);
// diagnostics switchboard for the bench / profiling scripts (never used by the product path)
extern "C" int d4_debug_set(const char* key, int value) {
    if (!key) return d4_fail("d4_debug_set: null key");
    if (!strcmp(key, "gemm_f16")) { d4_gemm_f16_debug(value); return 0; }
    if (!strcmp(key, "lp_fused")) { d4_lp_fused_debug(value); return 0; }          // 1: one CTA per frame, 2: persistent (default, large batches)
    return d4_fail("d4_debug_set: unknown key '%s'", key);
}
extern "C" int64_t d4_graph_replays(const d4_ctx* c) { return c ? c->graph_replays : 0; }
extern "C" int64_t d4_debug_get(const d4_ctx* c, const char* key) {
    if (!c || !key) return -1;
    if (!strcmp(key, "graph_enabled")) return c->use_graphs ? 1 : 0;
    if (!strcmp(key, "graph_keys")) return (int64_t)c->graphs.size();
    if (!strcmp(key, "graph_captured")) { int64_t n = 0; for (auto& kv : c->graphs) n += kv.second.exec != nullptr; return n; }
    if (!strcmp(key, "graph_capture_refused")) { int64_t n = 0; for (auto& kv : c->graphs) n += kv.second.direct; return n; }
    return -1;
}

#define D4_TRY(expr) do { int rc__ = (expr); if (rc__ != 0) return rc__; } while (0)

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// ------------------------------------------------------------------------------------------------ ctx
extern "C" int d4_ctx_create(const d4_config* cfg, d4_ctx** out) {
    if (!cfg || !out) return d4_fail("d4_ctx_create: null argument");
    d4_ctx* c = new d4_ctx();
    c->cfg = *cfg;
    c->D = cfg->dim; c->Dl = cfg->dim_latent; c->N = cfg->num_latent_tokens; c->nsp = cfg->num_spatial_tokens;
    c->nreg = cfg->num_register_tokens; c->L = cfg->depth; c->h = cfg->heads; c->hq = cfg->query_heads; c->d = cfg->dim_head;
    c->hp = cfg->pool_heads; c->dp = cfg->pool_dim_head; c->Dp = c->hp * c->dp;
    c->Dq = c->hq * c->d; c->Dkv = c->h * c->d;
    c->inner = cfg->ff_inner; c->inner_pad = cfg->ff_inner_pad;
    c->na = cfg->num_action_types; c->has_actions = c->na > 0;
    c->A_total = 0;
    for (int i = 0; i < c->na; ++i) { c->act_off[i] = c->A_total; c->A_total += cfg->action_sizes[i]; }
    c->S = 1 + c->nsp + c->nreg + c->has_actions + 1;
    c->same_len = (c->nsp == c->N);
    c->n_hid = 2 * c->L + 1;
    c->y = 0;
    for (int i = 0; i < c->L; ++i) { const int it = ((i + 1) % cfg->time_block_every) == 0; c->is_time.push_back(it); c->y += it; }
    c->NQ = c->Dq + 2 * c->Dkv + c->hq + c->h; c->ldq = round_up(c->NQ, 4);
    c->ldpq = round_up(c->Dp + c->hp, 4);
    c->ldfa = round_up(c->Dq + c->hq, 4);
    c->ldlog = round_up(c->A_total > 0 ? c->A_total : 1, 4);
    const char* bad = nullptr;
    if (c->hq % c->h != 0) bad = "query_heads must be a multiple of heads";
    else if (c->d % 4 != 0 || c->d > 128) bad = "dim_head must be a multiple of 4, <= 128";
    else if (c->D % 4 != 0) bad = "dim must be a multiple of 4";
    else if (c->S > 32) bad = "more than 32 tokens per frame unsupported";
    else if (c->N > 64 || c->nsp > 64) bad = "more than 64 latent / spatial tokens unsupported";
    else if (c->n_hid > 64) bad = "depth > 31 unsupported";
    else if (c->na > D4_MAX_ACTION_TYPES) bad = "too many action types";
    else if (c->inner_pad < c->inner || c->inner_pad % 4 != 0) bad = "ff_inner_pad must be >= ff_inner and a multiple of 4";
    else if (cfg->max_batch < 1 || cfg->max_time < 1) bad = "max_batch / max_time must be positive";
    else if (cfg->policy_layers > D4_MAX_MLP_LAYERS || cfg->value_layers > D4_MAX_MLP_LAYERS || cfg->terminal_layers > D4_MAX_MLP_LAYERS) bad = "MLP too deep";
    if (bad) { delete c; return d4_fail("d4_ctx_create: %s", bad); }
    { const char* f = getenv("D4_FUSE_POOLS"); c->fuse_pools = f ? atoi(f) != 0 : true; }
    { const char* f = getenv("D4_SPACE_MMA"); c->space_mma = f ? atoi(f) != 0 : true; }
    { const char* f = getenv("D4_FUSE_SS"); c->fuse_ss = f ? atoi(f) != 0 : true; }
    { const char* f = getenv("D4_GRAPH"); c->use_graphs = f ? atoi(f) != 0 : true; }
    { const char* f = getenv("D4_SKINNY"); c->skinny = f ? atoi(f) != 0 : true; }
    { const char* f = getenv("D4_GRAPH_MAX_ROWS"); if (f && atoi(f) > 0) c->graph_max_rows = atoi(f); }
    { const char* f = getenv("D4_TRIM_FINAL"); c->trim_final = f ? atoi(f) != 0 : true; }
    { const char* f = getenv("D4_TRIM_CONE"); c->trim_cone = f ? atoi(f) != 0 : true; }
    d4_engine_plan(c);
    *out = c;
    return 0;
}
static void drop_graphs(d4_ctx* c) {
    for (auto& kv : c->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    c->graphs.clear();
}
// the engine's own capture / replay stream and the two events that order it against the caller's stream (any stream, the legacy
// default stream included, which cannot itself be captured)
static int graph_stream(d4_ctx* c) {
    if (c->gstream) return 0;
    D4_CUDA_OK(cudaStreamCreateWithFlags(&c->gstream, cudaStreamNonBlocking));
    D4_CUDA_OK(cudaEventCreateWithFlags(&c->gev_in, cudaEventDisableTiming));
    D4_CUDA_OK(cudaEventCreateWithFlags(&c->gev_out, cudaEventDisableTiming));
    return 0;
}
extern "C" void d4_ctx_destroy(d4_ctx* ctx) {
    if (!ctx) return;
    drop_graphs(ctx);
    if (ctx->gstream) { cudaStreamSynchronize(ctx->gstream); cudaStreamDestroy(ctx->gstream); cudaEventDestroy(ctx->gev_in); cudaEventDestroy(ctx->gev_out); }
    for (auto& r : ctx->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto& r : ctx->prof_pool) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    delete ctx;
}

extern "C" int d4_set_weight(d4_ctx* c, const char* name, const float* p, int64_t numel) {
    if (!c || !name) return d4_fail("d4_set_weight: null argument");
    c->table[name] = std::make_pair(p, numel);
    c->bound = false;
    drop_graphs(c);           // captured frames hold the old pointers
    return 0;
}
extern "C" int d4_set_weight_scale(d4_ctx* c, const char* name, float scale) {
    if (!c || !name) return d4_fail("d4_set_weight_scale: null argument");
    if (!(scale > 0.f) || !(scale < 3.0e38f)) return d4_fail("d4_set_weight_scale: '%s' needs a positive finite scale", name);
    c->scales[name] = scale;
    c->bound = false;
    return 0;
}

int d4_engine_plan(d4_ctx* c) {
    const long long B = c->cfg.max_batch, M = B * c->S;
    long long off = 0;
    auto take = [&](long long nfloats) { long long o = off; off += (nfloats * 4 + 255) / 256 * 256; return o; };
    struct { float** p; long long n; } items[] = {
        {&c->b.lat_x, B * c->N * c->Dl}, {&c->b.lat_rstd, B * c->N}, {&c->b.kv_l, B * c->N * 2 * c->Dkv}, {&c->b.att_l, B * c->nsp * c->Dq},
        {&c->b.hid, (long long)c->n_hid * M * c->D}, {&c->b.hid_rstd, (long long)c->n_hid * M}, {&c->b.x_cur, M * c->D}, {&c->b.x_rstd, (long long)(c->L + 1) * M},
        {&c->b.qkvgm, M * c->ldq}, {&c->b.v0, M * c->Dkv}, {&c->b.attn_o, M * c->Dq}, {&c->b.ff_mid, M * c->inner_pad},
        {&c->b.pool_qg, M * c->ldpq}, {&c->b.pool_kv, (long long)c->n_hid * M * 2 * c->Dp}, {&c->b.pool_att, M * c->Dp},
        {&c->b.fa_kv, M * 2 * c->Dkv}, {&c->b.fa_q, B * c->ldfa}, {&c->b.fa_att, B * c->Dq}, {&c->b.ag_rstd, B},
        {&c->b.sp_n, B * c->nsp * c->D}, {&c->b.sp_n2, B * c->nsp * c->D}, {&c->b.sp_kv, B * c->nsp * 2 * c->Dkv},
        {&c->b.lp_att, B * c->N * c->Dq}, {&c->b.pred, B * c->N * c->Dl},
        {&c->b.hbuf0, B * (long long)std::max(std::max(c->cfg.policy_hidden, c->cfg.value_hidden), std::max(c->cfg.terminal_hidden, 4))},
        {&c->b.hbuf1, B * (long long)std::max(std::max(c->cfg.policy_hidden, c->cfg.value_hidden), std::max(c->cfg.terminal_hidden, 4))},
        {&c->b.logits, B * c->ldlog}, {&c->b.bins, B * (long long)std::max(std::max(c->cfg.reward_bins, c->cfg.value_bins), 4)},
        {&c->b.agent, B * c->D}, {&c->b.term_in, B * c->Dl},
        {&c->b.fin_x, B * (long long)std::max(c->nsp, 1) * c->D}, {&c->b.fin_rstd, B * (long long)std::max(c->nsp, 1)},
        {&c->b.fin_ctx_rs, (long long)c->n_hid * B * std::max(c->nsp, 1)}, {&c->b.fin_rstd2, B * (long long)std::max(c->nsp, 1)},
    };
    for (auto& it : items) *it.p = reinterpret_cast<float*>(take(it.n));   // offsets for now; rebased in d4_set_buffers
    c->b.sizes_offs = reinterpret_cast<int*>(take(2 * D4_MAX_ACTION_TYPES));
    if (c->tf_mode) {
        c->tfb.fa_q = reinterpret_cast<float*>(take(B * c->tf_ns * c->ldfa)); c->tfb.fa_att = reinterpret_cast<float*>(take(B * c->tf_ns * c->Dq));
        c->tfb.sp_rstd = reinterpret_cast<float*>(take(B * c->tf_ns));
    }
    if (c->use_graphs) {      // dense staging rows of one frame's inputs / outputs (offsets for now, like the rest)
        const long long A = std::max(c->A_total, 1), na = std::max(c->na, 1);
        auto& g = c->gio;
        g.noise = reinterpret_cast<float*>(take(B * c->N * c->Dl)); g.latents = reinterpret_cast<float*>(take(B * c->N * c->Dl));
        g.act_u = reinterpret_cast<float*>(take(B * A)); g.logits = reinterpret_cast<float*>(take(B * A));
        g.term_u = reinterpret_cast<float*>(take(B)); g.rewards = reinterpret_cast<float*>(take(B)); g.values = reinterpret_cast<float*>(take(B));
        g.agent = reinterpret_cast<float*>(take(B * c->D)); g.logp = reinterpret_cast<float*>(take(B * na));
        g.prev_actions = reinterpret_cast<long long*>(take(2 * B * na)); g.actions = reinterpret_cast<long long*>(take(2 * B * na));
        g.tasks = reinterpret_cast<long long*>(take(2 * B)); g.lens = reinterpret_cast<long long*>(take(2 * B));
        g.terminals = reinterpret_cast<unsigned char*>(take((B + 3) / 4));
    }
    c->ws_need = off;
    return 0;
}

extern "C" int64_t d4_workspace_bytes(const d4_ctx* c) { return c ? c->ws_need : 0; }
extern "C" int64_t d4_kv_bytes(const d4_ctx* c) {
    if (!c) return 0;
    return (int64_t)std::max(c->y, 1) * 2 * c->cfg.max_batch * c->S * c->h * (int64_t)c->cfg.max_time * c->d * 4;
}

extern "C" int d4_set_buffers(d4_ctx* c, void* workspace, int64_t workspace_bytes, float* kv, int64_t kv_bytes) {
    if (!c) return d4_fail("d4_set_buffers: null ctx");
    if (workspace_bytes < c->ws_need) return d4_fail("d4_set_buffers: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)c->ws_need);
    if (kv_bytes < d4_kv_bytes(c)) return d4_fail("d4_set_buffers: kv buffer %lld < %lld bytes", (long long)kv_bytes, (long long)d4_kv_bytes(c));
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) || (reinterpret_cast<uintptr_t>(kv) & 15)) return d4_fail("d4_set_buffers: buffers must be 256-byte aligned");
    if (c->ws) return d4_fail("d4_set_buffers: buffers already set");
    d4_engine_plan(c);
    unsigned char* base = static_cast<unsigned char*>(workspace);
    float** ptrs = reinterpret_cast<float**>(&c->b);
    const int nptr = (int)(offsetof(decltype(c->b), sizes_offs) / sizeof(float*));
    for (int i = 0; i < nptr; ++i) ptrs[i] = reinterpret_cast<float*>(base + reinterpret_cast<uintptr_t>(ptrs[i]));
    c->b.sizes_offs = reinterpret_cast<int*>(base + reinterpret_cast<uintptr_t>(c->b.sizes_offs));
    if (c->tf_mode) {
        void** tp = reinterpret_cast<void**>(&c->tfb);
        for (size_t i = 0; i < sizeof(c->tfb) / sizeof(void*); ++i) tp[i] = base + reinterpret_cast<uintptr_t>(tp[i]);
    }
    if (c->use_graphs) {
        void** gp = reinterpret_cast<void**>(&c->gio);
        for (size_t i = 0; i < sizeof(c->gio) / sizeof(void*); ++i) gp[i] = base + reinterpret_cast<uintptr_t>(gp[i]);
    }
    drop_graphs(c);
    c->ws = base; c->ws_bytes = workspace_bytes; c->kv = kv; c->kv_bytes_ = kv_bytes;
    // ff_mid pad columns must read as zero (they meet zero-padded weight columns)
    D4_CUDA_OK(cudaMemset(c->b.ff_mid, 0, (size_t)c->cfg.max_batch * c->S * c->inner_pad * 4));
    int so[2 * D4_MAX_ACTION_TYPES] = {0};
    for (int i = 0; i < c->na; ++i) { so[i] = c->cfg.action_sizes[i]; so[c->na + i] = c->act_off[i]; }
    D4_CUDA_OK(cudaMemcpy(c->b.sizes_offs, so, sizeof(so), cudaMemcpyHostToDevice));
    return 0;
}

// ------------------------------------------------------------------------------------------------ binding
namespace {
struct Binder {
    d4_ctx* c; const char* missing = nullptr; std::string miss_store;
    const float* get(const std::string& name, int64_t numel, bool optional = false) {
        auto it = c->table.find(name);
        if (it == c->table.end() || it->second.first == nullptr) {
            if (!optional && !missing) { miss_store = name; missing = miss_store.c_str(); }
            return nullptr;
        }
        if (numel >= 0 && it->second.second != numel && !missing) {
            miss_store = name + " (numel " + std::to_string(it->second.second) + " != expected " + std::to_string(numel) + ")";
            missing = miss_store.c_str();
        }
        return it->second.first;
    }
    LinW lin(const std::string& name, int64_t numel) {
        LinW w; w.w = get(name, numel);
        const bool optional = !d4_prec_split(c->cfg.precision);
        w.hi = get(name + ".hi", numel, optional);
        w.lo = get(name + ".lo", numel, optional);
        if (c->cfg.precision == D4_PREC_F16X3) {
            // fp16 words of the pre-scaled weight (same element count, 2 bytes each) and 1 / q; a weight registered without
            // them simply stays on the 3xTF32 kernel
            w.h_hi = get(name + ".h16hi", numel, true);
            w.h_lo = get(name + ".h16lo", numel, true);
            auto it = c->scales.find(name);
            if (w.h_hi && w.h_lo && it != c->scales.end()) w.h_scale = it->second;
            else w.h_hi = w.h_lo = nullptr;
        }
        return w;
    }
};
}  // namespace

static int tf_bind(d4_ctx* c);
extern "C" int d4_bind(d4_ctx* c) {
    if (!c) return d4_fail("d4_bind: null ctx");
    if (c->tf_mode) return tf_bind(c);
    Binder B{c};
    const int D = c->D, Dl = c->Dl, Dq = c->Dq, Dkv = c->Dkv, h = c->h, hq = c->hq, d = c->d, Dp = c->Dp, hp = c->hp;
    const int half = D / 2;
    c->sig_emb = B.get("sig_emb", (int64_t)c->cfg.max_steps * half);
    c->step_emb = B.get("step_emb", -1);
    c->registers = B.get("registers", (int64_t)c->nreg * D, c->nreg == 0);
    c->agent_embed = B.get("agent_embed", D);
    c->action_learned = B.get("action_learned", D, !c->has_actions);
    c->action_emb = B.get("action_emb", (int64_t)c->A_total * D, !c->has_actions);
    c->task_emb = B.get("task_emb", -1, true);
    if (c->same_len) {
        c->l2s_w = B.lin("l2s.w", (int64_t)D * Dl); c->l2s_b = B.get("l2s.b", D);
        c->lp_w = B.lin("lp.w", (int64_t)Dl * D);
    } else {
        c->l2s_w_kv = B.lin("l2s.w_kv", (int64_t)2 * Dkv * Dl); c->l2s_w_out = B.lin("l2s.w_out", (int64_t)D * Dq);
        c->l2s_q = B.get("l2s.q", (int64_t)c->nsp * Dq); c->l2s_gate = B.get("l2s.gate", (int64_t)c->nsp * hq);
        c->l2s_k_gamma = B.get("l2s.k_gamma", (int64_t)h * d);
        c->lp_norm_ctx = B.get("lp.norm_ctx", D); c->lp_w_kv = B.lin("lp.w_kv", (int64_t)2 * Dkv * D);
        c->lp_q = B.get("lp.q", (int64_t)c->N * Dq); c->lp_gate = B.get("lp.gate", (int64_t)c->N * hq);
        c->lp_k_gamma = B.get("lp.k_gamma", (int64_t)h * d); c->lp_w_comb = B.lin("lp.w_comb", (int64_t)Dl * Dq);
    }
    c->lp_norm0 = B.get("lp.norm0", D);
    c->vr_w = B.lin("vr.w", (int64_t)Dkv * D);
    c->inv_freq = B.get("inv_freq", d / 2);
    auto bind_ff = [&](const std::string& p, FFW& f) {
        f.w_in = B.lin(p + ".w_in", (int64_t)2 * c->inner * D); f.b_in = B.get(p + ".b_in", 2 * c->inner);
        f.w_out = B.lin(p + ".w_out", (int64_t)D * c->inner_pad); f.b_out = B.get(p + ".b_out", D);
    };
    auto bind_pool = [&](const std::string& p, PoolW& w) {
        w.w_qg = B.lin(p + ".w_qg", (int64_t)(Dp + hp) * D); w.w_kv = B.lin(p + ".w_kv", (int64_t)2 * Dp * D);
        w.k_gamma = B.get(p + ".k_gamma", (int64_t)hp * c->dp); w.w_out = B.lin(p + ".w_out", (int64_t)D * Dp);
    };
    c->attn.assign(c->L, AttnLayerW()); c->ff.assign(c->L, FFW()); c->pools.assign(c->L > 0 ? c->L - 1 : 0, PoolW());
    for (int i = 0; i < c->L; ++i) {
        const std::string p = "L" + std::to_string(i);
        c->attn[i].w = B.lin(p + ".attn.w", (int64_t)c->NQ * D); c->attn[i].b = B.get(p + ".attn.b", c->NQ);
        c->attn[i].k_gamma = B.get(p + ".attn.k_gamma", (int64_t)h * d); c->attn[i].w_out = B.lin(p + ".attn.w_out", (int64_t)D * Dq);
        bind_ff(p + ".ff", c->ff[i]);
        if (i != c->L - 1) bind_pool("P" + std::to_string(i), c->pools[i]);
    }
    bind_pool("PF", c->pool_final);
    c->fa.w_qg = B.lin("FA.w_qg", (int64_t)(Dq + hq) * D); c->fa.w_kv = B.lin("FA.w_kv", (int64_t)2 * Dkv * D);
    c->fa.k_gamma = B.get("FA.k_gamma", (int64_t)h * d); c->fa.w_out = B.lin("FA.w_out", (int64_t)D * Dq);
    bind_ff("FAFF", c->fa_ff);
    c->reward_w = B.get("reward.w", (int64_t)c->cfg.reward_bins * D);
    c->reward_centers = B.get("reward.centers", c->cfg.reward_bins);
    auto bind_mlp = [&](const std::string& p, MlpW& m, int layers, int din, int hidden, int dout) {
        m.layers = layers;
        for (int l = 0; l <= layers; ++l) m.dims[l] = (l == 0) ? din : (l == layers ? dout : hidden);
        for (int l = 0; l < layers; ++l) {
            const std::string q = p + "." + std::to_string(l);
            m.w[l] = B.get(q + ".w", (int64_t)m.dims[l + 1] * m.dims[l]); m.b[l] = B.get(q + ".b", m.dims[l + 1]);
            m.hi[l] = B.get(q + ".w.hi", (int64_t)m.dims[l + 1] * m.dims[l], true); m.lo[l] = B.get(q + ".w.lo", (int64_t)m.dims[l + 1] * m.dims[l], true);
            m.h_hi[l] = m.h_lo[l] = nullptr; m.h_scale[l] = 1.f;
            if (c->cfg.precision == D4_PREC_F16X3) {          // the rollout's head MLPs on the fp16 split GEMM when the caller registered the words + scale
                const void* hh = B.get(q + ".w.h16hi", (int64_t)m.dims[l + 1] * m.dims[l], true);
                const void* hl = B.get(q + ".w.h16lo", (int64_t)m.dims[l + 1] * m.dims[l], true);
                auto it = c->scales.find(q + ".w");
                if (hh && hl && it != c->scales.end()) { m.h_hi[l] = hh; m.h_lo[l] = hl; m.h_scale[l] = it->second; }
            }
            m.wthi[l] = B.get(q + ".wt.hi", (int64_t)m.dims[l + 1] * m.dims[l], true); m.wtlo[l] = B.get(q + ".wt.lo", (int64_t)m.dims[l + 1] * m.dims[l], true);
            if (l < layers - 1) { m.lnw[l] = B.get(q + ".lnw", m.dims[l + 1]); m.lnb[l] = B.get(q + ".lnb", m.dims[l + 1]); }
            else { m.lnw[l] = m.lnb[l] = nullptr; }
        }
    };
    if (c->has_actions) {
        bind_mlp("policy", c->policy, c->cfg.policy_layers, D, c->cfg.policy_hidden, c->cfg.policy_hidden);
        bind_mlp("value", c->value, c->cfg.value_layers, D, c->cfg.value_hidden, c->cfg.value_bins);
        c->value_centers = B.get("value.centers", c->cfg.value_bins);
        auto it = c->table.find("unembed");
        if (it == c->table.end()) { if (!B.missing) B.missing = "unembed"; }
        else { c->unembed = it->second.first; c->unembed_ld = it->second.second; }   // numel slot carries the row stride
    }
    if (c->cfg.predict_terminals) bind_mlp("terminal", c->terminal, c->cfg.terminal_layers, Dl, c->cfg.terminal_hidden, 1);
    if (B.missing) return d4_fail("d4_bind: weight '%s' missing or mis-sized", B.missing);
    c->bound = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------ profiling
int d4_prof_begin(d4_ctx* c, int cls, double work, cudaStream_t s) {
    if (!c->prof_on) return -1;
    d4_ctx::ProfRec r;
    if (!c->prof_pool.empty()) { r = c->prof_pool.back(); c->prof_pool.pop_back(); }
    else {
        if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return -1;
    }
    r.cls = cls; r.work = work;
    cudaEventRecord(r.a, s);
    c->prof.push_back(r);
    return (int)c->prof.size() - 1;
}
void d4_prof_end(d4_ctx* c, int handle, cudaStream_t s) {
    if (handle >= 0) cudaEventRecord(c->prof[handle].b, s);
}
extern "C" int d4_profile(d4_ctx* c, int enable) {
    if (!c) return d4_fail("d4_profile: null ctx");
    c->prof_on = enable != 0;
    return 0;
}
extern "C" int d4_profile_read(d4_ctx* c, double* out) {
    if (!c || !out) return d4_fail("d4_profile_read: null argument");
    for (int i = 0; i < D4_PROF_CLASSES * 3; ++i) out[i] = 0.0;
    for (auto& r : c->prof) {
        D4_CUDA_OK(cudaEventSynchronize(r.b));
        float ms = 0.f;
        D4_CUDA_OK(cudaEventElapsedTime(&ms, r.a, r.b));
        out[r.cls * 3 + 0] += ms; out[r.cls * 3 + 1] += 1.0; out[r.cls * 3 + 2] += r.work;
        c->prof_pool.push_back(r);
    }
    c->prof.clear();
    return 0;
}

// ------------------------------------------------------------------------------------------------ GEMM dispatch
int d4_engine_gemm(d4_ctx* c, GemmArgs g, const LinW& w, int force_fp32, cudaStream_t s) {
    g.W = w.w;
    const int prec = force_fp32 ? D4_PREC_FP32 : c->cfg.precision;
    const int ph = d4_prof_begin(c, D4_CLS_GEMM, 2.0 * g.M * g.N * g.K, s);
    int rc;
    if (c->skinny && d4_gemm_skinny_supported(g)) {
        // a handful of rows (tiny batches; the B-row projections and heads of any batch <= 32): a weight stream, exact fp32
        rc = d4_gemm_skinny(g, s);
    } else if (prec == D4_PREC_F16X3 && w.h_hi && w.h_lo && d4_gemm_f16x3_supported(g, w.h_hi, w.h_lo)) {
        // fp16 3-term split on kind::f16 (gemm_f16.cu); the fp16 arrays share the fp32 weight's (N, ldw) shape
        g.W = static_cast<const float*>(w.h_hi); g.W_lo = static_cast<const float*>(w.h_lo);
        rc = d4_gemm_f16x3(g, w.h_scale, 0, s);
    } else if (prec != D4_PREC_FP32 && d4_gemm_tc_supported(g)) {
        if (d4_prec_split(prec) && w.hi && w.lo) { g.W = w.hi; g.W_lo = w.lo; rc = d4_gemm_tc(g, 3, s); }
        else rc = d4_gemm_tc(g, 1, s);
    } else {
        rc = d4_gemm_simt(g, s);
    }
    d4_prof_end(c, ph, s);
    return rc;
}

static inline LinW plain(const float* w) { LinW l; l.w = w; return l; }

// x-mlps normed MLP forward (Linear -> LayerNorm -> SiLU)* -> Linear.  Exact fp32 FMA unless the caller registered a tf32
// hi/lo split of the layer's weight and the engine runs in tf32x3 (fp32-accurate 3-term tensor-core product).
int d4_mlp_forward(d4_ctx* c, const MlpW& mlp, const float* x, long long ldx, int M, float* buf0, float* buf1, float* out, long long ldo,
                   int allow_tensor, cudaStream_t s) {
    const float* cur = x; long long ldc = ldx;
    for (int l = 0; l < mlp.layers; ++l) {
        const bool last = (l == mlp.layers - 1);
        float* dst = last ? out : ((l & 1) ? buf1 : buf0);
        const long long ldd = last ? ldo : mlp.dims[l + 1];
        GemmArgs g = gemm_args(cur, ldc, mlp.w[l], mlp.dims[l], dst, ldd, M, mlp.dims[l + 1], mlp.dims[l]);
        g.bias = mlp.b[l];
        LinW lw; lw.w = mlp.w[l]; lw.hi = mlp.hi[l]; lw.lo = mlp.lo[l];
        if (allow_tensor) { lw.h_hi = mlp.h_hi[l]; lw.h_lo = mlp.h_lo[l]; lw.h_scale = mlp.h_scale[l]; }
        const int exact = !(allow_tensor && d4_prec_split(c->cfg.precision) && lw.hi && lw.lo);
        D4_TRY(d4_engine_gemm(c, g, lw, exact, s));
        if (!last) D4_TRY(d4_ln_act_rows(dst, ldd, mlp.lnw[l], mlp.lnb[l], M, mlp.dims[l + 1], dst, ldd, D4_ACT_SILU, nullptr, nullptr, s));
        cur = dst; ldc = ldd;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------ the pass
namespace {

// xq_ss / ctx_ss: the row statistic buffers hold sums of squares (fused into the producing GEMMs) instead of rstd;
// ss_out: where the output rows' sums of squares are accumulated (or nullptr)
int run_pool(d4_ctx* c, const PoolW& P, int M, const float* xq, const float* xq_rstd, int xq_ss, int ctx_ss, int n, float* out, float* ss_out,
             cudaStream_t s) {
    const int D = c->D, Dp = c->Dp, hp = c->hp, dp = c->dp;
    SmallAttnArgs a; memset(&a, 0, sizeof(a));
    a.nb = M; a.hkv = hp; a.g = 1; a.d = dp; a.nq = 1; a.n = n;
    a.q = c->b.pool_qg; a.q_sb = c->ldpq; a.q_si = 0;
    a.k = c->b.pool_kv; a.k_sb = 2 * Dp; a.k_sj = (long long)M * 2 * Dp;
    a.v = c->b.pool_kv + Dp; a.v_sb = a.k_sb; a.v_sj = a.k_sj;
    a.k_gamma = P.k_gamma;
    a.out = c->b.pool_att; a.out_sb = Dp; a.out_si = 0;
    a.scale = 1.f / sqrtf((float)dp);
    // the 4 gate rows ride along in w_qg (rows Dp .. Dp+hp); when the pool kernel can form the gate logits itself the q
    // projection is an exact 256-column GEMM (one full tile instead of 260 -> 3 x 128 with 48 % padding)
    a.gate_x = xq; a.gate_x_ld = D; a.gate_rstd = xq_rstd; a.gate_w = P.w_qg.w + (long long)Dp * D; a.gate_D = D; a.gate_rstd_is_ss = xq_ss;
    const bool gate_in_kernel = d4_pool_attn_ok(a) != 0;
    if (!gate_in_kernel) {
        a.gate_x = nullptr; a.gate_rstd = nullptr; a.gate_w = nullptr; a.gate_D = 0; a.gate_rstd_is_ss = 0;
        a.gate = c->b.pool_qg + Dp; a.gate_sb = c->ldpq; a.gate_si = 0;
    }
    {   // query (+ gate logits) from the normed token
        GemmArgs g = gemm_args(xq, D, nullptr, D, c->b.pool_qg, c->ldpq, M, gate_in_kernel ? Dp : Dp + hp, D);
        g.row_scale = xq_rstd; g.rs_mode = xq_ss;
        D4_TRY(d4_engine_gemm(c, g, P.w_qg, 0, s));
    }
    {   // keys / values of all hiddens so far, re-projected by this pool (reference dreamer4.py:2164-2177)
        GemmArgs g = gemm_args(c->b.hid, D, nullptr, D, c->b.pool_kv, 2 * Dp, n * M, 2 * Dp, D);
        g.row_scale = c->b.hid_rstd; g.rs_mode = ctx_ss;
        D4_TRY(d4_engine_gemm(c, g, P.w_kv, 0, s));
    }
    { const int ph = d4_prof_begin(c, D4_CLS_SMALL_ATTN, 0.0, s); const int rc = d4_small_attn(a, s); d4_prof_end(c, ph, s); D4_TRY(rc); }
    GemmArgs g = gemm_args(c->b.pool_att, Dp, nullptr, Dp, out, D, M, D, Dp);
    g.residual = xq; g.ldr = D; g.ss_out = ss_out;
    return d4_engine_gemm(c, g, P.w_out, 0, s);
}

// run_pool for a SUBSET of token rows: the pool is per token (each token attends over its own history of hiddens), so a pass that
// only reads some rows of the result - the 4 spatial tokens of every frame on a denoise pass, the agent token on the clean pass -
// needs the keys / values of those rows only.  Rows = groups of `grp` tokens at offset `goff` of every frame; the context rows are read
// from the snapshots through that row map.  mapped = 0: xq and out hold the B * grp rows compactly; mapped = 1: xq and out are (B, S, D)
// tensors and the rows sit at their place in them (the layers that run on a row subset, see run_pass).
int run_pool_rows(d4_ctx* c, const PoolW& P, int B, int grp, int goff, const float* xq, int mapped, int ctx_ss, int n, float* out, cudaStream_t s) {
    const int D = c->D, Dp = c->Dp, hp = c->hp, dp = c->dp, S = c->S;
    const int Mq = B * grp;
    const RowMap rows = rowmap(grp, S, goff);
    const RowMap io = mapped ? rows : rowmap_identity();
    D4_TRY(d4_row_rstd(xq, D, io, Mq, D, c->b.fin_rstd, s));
    SmallAttnArgs a; memset(&a, 0, sizeof(a));
    a.nb = Mq; a.hkv = hp; a.g = 1; a.d = dp; a.nq = 1; a.n = n;
    a.q = c->b.pool_qg; a.q_sb = c->ldpq; a.q_si = 0;
    a.k = c->b.pool_kv; a.k_sb = 2 * Dp; a.k_sj = (long long)Mq * 2 * Dp;
    a.v = c->b.pool_kv + Dp; a.v_sb = a.k_sb; a.v_sj = a.k_sj;
    a.k_gamma = P.k_gamma;
    a.out = c->b.pool_att; a.out_sb = Dp; a.out_si = 0;
    a.scale = 1.f / sqrtf((float)dp);
    a.gate_x = xq; a.gate_x_ld = D; a.gate_rstd = c->b.fin_rstd; a.gate_w = P.w_qg.w + (long long)Dp * D; a.gate_D = D; a.gate_rstd_is_ss = 0;
    const bool gate_in_kernel = !mapped && d4_pool_attn_ok(a) != 0;       // the in-kernel gate reads the query rows compactly
    if (!gate_in_kernel) {
        a.gate_x = nullptr; a.gate_rstd = nullptr; a.gate_w = nullptr; a.gate_D = 0;
        a.gate = c->b.pool_qg + Dp; a.gate_sb = c->ldpq; a.gate_si = 0;
    }
    {
        GemmArgs g = gemm_args(xq, D, nullptr, D, c->b.pool_qg, c->ldpq, Mq, gate_in_kernel ? Dp : Dp + hp, D);
        g.amap = io; g.row_scale = c->b.fin_rstd;
        D4_TRY(d4_engine_gemm(c, g, P.w_qg, 0, s));
    }
    {   // the statistics of the context rows, gathered next to their compact GEMM rows (row (j, b, i) of n x B x grp)
        D4_TRY(d4_gather_rows(c->b.hid_rstd, 1, rows, (long long)n * Mq, 1, c->b.fin_ctx_rs, 1, s));
        GemmArgs g = gemm_args(c->b.hid, D, nullptr, D, c->b.pool_kv, 2 * Dp, n * Mq, 2 * Dp, D);
        g.amap = rows; g.row_scale = c->b.fin_ctx_rs; g.rs_mode = ctx_ss;
        D4_TRY(d4_engine_gemm(c, g, P.w_kv, 0, s));
    }
    { const int ph = d4_prof_begin(c, D4_CLS_SMALL_ATTN, 0.0, s); const int rc = d4_small_attn(a, s); d4_prof_end(c, ph, s); D4_TRY(rc); }
    GemmArgs g = gemm_args(c->b.pool_att, Dp, nullptr, Dp, out, D, Mq, D, Dp);
    g.residual = xq; g.ldr = D; g.cmap = io;          // residual rows follow the output row map
    return d4_engine_gemm(c, g, P.w_out, 0, s);
}

int run_ff(d4_ctx* c, const FFW& F, int M, const float* x, long long ldx, RowMap xmap, const float* x_rstd, int x_ss, float* out, long long ldo, RowMap omap,
           float* ss_out, cudaStream_t s) {
    const int D = c->D;
    GemmArgs g = gemm_args(x, ldx, nullptr, D, c->b.ff_mid, c->inner_pad, M, 2 * c->inner, D);
    g.amap = xmap; g.row_scale = x_rstd; g.rs_mode = x_ss; g.bias = F.b_in; g.act = c->cfg.ff_act == 1 ? D4_ACT_GLU_GELU : D4_ACT_GLU_SILU;
    D4_TRY(d4_engine_gemm(c, g, F.w_in, 0, s));
    GemmArgs g2 = gemm_args(c->b.ff_mid, c->inner_pad, nullptr, c->inner_pad, out, ldo, M, D, c->inner_pad);
    g2.bias = F.b_out; g2.residual = x; g2.ldr = ldx; g2.cmap = omap; g2.ss_out = ss_out;
    // residual rows follow the output row map (x and out share their row layout in every use)
    return d4_engine_gemm(c, g2, F.w_out, 0, s);
}

// to_latent_pred after its first RMSNorm (the normed spatial tokens are in sp_n): Linear, or learned-query pool + Linear
// (reference dreamer4.py:4830-4834, 7251)
int finish_latent_pred(d4_ctx* c, int B, float* pred_out, cudaStream_t s) {
    const int D = c->D, Dl = c->Dl, N = c->N, nsp = c->nsp, Dq = c->Dq, Dkv = c->Dkv, h = c->h, hq = c->hq, d = c->d;
    const float att_scale = 1.f / sqrtf((float)d);
    if (c->same_len) {
        GemmArgs g = gemm_args(c->b.sp_n, D, nullptr, D, pred_out, Dl, B * N, Dl, D);
        return d4_engine_gemm(c, g, c->lp_w, 0, s);
    }
    D4_TRY(d4_rmsnorm_rows(c->b.sp_n, D, rowmap_identity(), c->lp_norm_ctx, B * nsp, D, c->b.sp_n2, D, s));
    GemmArgs g = gemm_args(c->b.sp_n2, D, nullptr, D, c->b.sp_kv, 2 * Dkv, B * nsp, 2 * Dkv, D);
    D4_TRY(d4_engine_gemm(c, g, c->lp_w_kv, 0, s));
    LpArgs pa; memset(&pa, 0, sizeof(pa));
    pa.B = B; pa.N = N; pa.Dl = Dl; pa.nsp = nsp; pa.h = h; pa.hq = hq; pa.d = d;
    pa.kv = c->b.sp_kv; pa.q = c->lp_q; pa.gate = c->lp_gate; pa.k_gamma = c->lp_k_gamma; pa.w_comb = c->lp_w_comb.w;
    pa.pred = pred_out; pa.scale = att_scale;
    if (c->fuse_pools && d4_lp_fused_supported(pa)) {
        const int ph = d4_prof_begin(c, D4_CLS_SMALL_ATTN, 0.0, s); const int rc = d4_lp_fused(pa, s); d4_prof_end(c, ph, s);
        return rc;
    }
    SmallAttnArgs a; memset(&a, 0, sizeof(a));
    a.nb = B; a.hkv = h; a.g = hq / h; a.d = d; a.nq = N; a.n = nsp;
    a.q = c->lp_q; a.q_sb = 0; a.q_si = Dq;
    a.k = c->b.sp_kv; a.k_sb = (long long)nsp * 2 * Dkv; a.k_sj = 2 * Dkv;
    a.v = c->b.sp_kv + Dkv; a.v_sb = a.k_sb; a.v_sj = a.k_sj;
    a.k_gamma = c->lp_k_gamma;
    a.gate = c->lp_gate; a.gate_sb = 0; a.gate_si = hq;
    a.out = c->b.lp_att; a.out_sb = (long long)N * Dq; a.out_si = Dq;
    a.scale = att_scale;
    { const int ph = d4_prof_begin(c, D4_CLS_SMALL_ATTN, 0.0, s); const int rc = d4_small_attn(a, s); d4_prof_end(c, ph, s); D4_TRY(rc); }
    GemmArgs g2 = gemm_args(c->b.lp_att, Dq, nullptr, Dq, pred_out, Dl, B * N, Dl, Dq);
    return d4_engine_gemm(c, g2, c->lp_w_comb, 0, s);
}

int run_pass(d4_ctx* c, int B, const float* latent, int signal, int step_log2, const int64_t* prev_actions, int64_t pa_stride,
             const int64_t* tasks, int t, int commit, float* pred_out, float* agent_out, cudaStream_t s,
             const int64_t* signal_rows = nullptr, const int64_t* step_rows = nullptr) {
    const int S = c->S, D = c->D, Dl = c->Dl, N = c->N, nsp = c->nsp, Dq = c->Dq, Dkv = c->Dkv, h = c->h, hq = c->hq, d = c->d, L = c->L;
    const int M = B * S;
    const long long MD = (long long)M * D;
    auto hid = [&](int j) { return c->b.hid + (long long)j * MD; };
    auto hrs = [&](int j) { return c->b.hid_rstd + (long long)j * M; };
    const float att_scale = 1.f / sqrtf((float)d);

    // ---- tokens of the new frame (reference dreamer4.py:7168-7222)
    if (c->same_len) {
        GemmArgs g = gemm_args(latent, Dl, nullptr, Dl, hid(0), D, B * N, D, Dl);
        g.bias = c->l2s_b; g.cmap = rowmap(nsp, S, 1);
        D4_TRY(d4_engine_gemm(c, g, c->l2s_w, 0, s));
    } else {
        L2sArgs la; memset(&la, 0, sizeof(la));
        la.B = B; la.N = N; la.Dl = Dl; la.nsp = nsp; la.h = h; la.hq = hq; la.g = hq / h; la.d = d; la.Dq = Dq;
        la.latent = latent; la.w_k = c->l2s_w_kv.w; la.w_v = c->l2s_w_kv.w + (long long)Dkv * Dl;
        la.q = c->l2s_q; la.gate = c->l2s_gate; la.k_gamma = c->l2s_k_gamma; la.out = c->b.att_l; la.scale = att_scale;
        la.allow_tensor = (c->cfg.precision != D4_PREC_FP32) && c->space_mma;
        if (c->fuse_pools && d4_l2s_fused_supported(la)) {
            const int ph = d4_prof_begin(c, D4_CLS_SMALL_ATTN, 0.0, s); const int rc = d4_l2s_fused(la, s); d4_prof_end(c, ph, s); D4_TRY(rc);
        } else {
            D4_TRY(d4_row_rstd(latent, Dl, rowmap_identity(), B * N, Dl, c->b.lat_rstd, s));
            GemmArgs g = gemm_args(latent, Dl, nullptr, Dl, c->b.kv_l, 2 * Dkv, B * N, 2 * Dkv, Dl);
            g.row_scale = c->b.lat_rstd;
            D4_TRY(d4_engine_gemm(c, g, c->l2s_w_kv, 0, s));
            SmallAttnArgs a; memset(&a, 0, sizeof(a));
            a.nb = B; a.hkv = h; a.g = hq / h; a.d = d; a.nq = nsp; a.n = N;
            a.q = c->l2s_q; a.q_sb = 0; a.q_si = Dq;
            a.k = c->b.kv_l; a.k_sb = (long long)N * 2 * Dkv; a.k_sj = 2 * Dkv;
            a.v = c->b.kv_l + Dkv; a.v_sb = a.k_sb; a.v_sj = a.k_sj;
            a.k_gamma = c->l2s_k_gamma;
            a.gate = c->l2s_gate; a.gate_sb = 0; a.gate_si = hq;
            a.out = c->b.att_l; a.out_sb = (long long)nsp * Dq; a.out_si = Dq;
            a.scale = att_scale;
            { const int ph = d4_prof_begin(c, D4_CLS_SMALL_ATTN, 0.0, s); const int rc = d4_small_attn(a, s); d4_prof_end(c, ph, s); D4_TRY(rc); }
        }
        GemmArgs g2 = gemm_args(c->b.att_l, Dq, nullptr, Dq, hid(0), D, B * nsp, D, Dq);
        g2.cmap = rowmap(nsp, S, 1);
        D4_TRY(d4_engine_gemm(c, g2, c->l2s_w_out, 0, s));
    }
    {
        AssembleArgs a; memset(&a, 0, sizeof(a));
        a.tokens = hid(0); a.B = B; a.S = S; a.D = D; a.nsp = nsp; a.nreg = c->nreg; a.has_actions = c->has_actions; a.na = c->na;
        a.sig_emb = c->sig_emb; a.step_emb = c->step_emb; a.signal = signal; a.step = step_log2;
        a.registers = c->registers; a.agent_embed = c->agent_embed; a.action_learned = c->action_learned; a.action_emb = c->action_emb;
        a.prev_actions = reinterpret_cast<const long long*>(prev_actions); a.pa_stride = pa_stride;
        for (int i = 0; i < c->na; ++i) a.act_off[i] = c->act_off[i];
        a.task_emb = c->task_emb; a.tasks = (tasks && c->task_emb) ? reinterpret_cast<const long long*>(tasks) : nullptr;
        a.signal_rows = reinterpret_cast<const long long*>(signal_rows); a.step_rows = reinterpret_cast<const long long*>(step_rows);
        D4_TRY(d4_assemble_tokens(a, s));
    }
    // RMS statistics.  Fused mode (CTA-pair tensor-core GEMMs, D <= 512 so a row spans at most two column tiles and the two
    // atomic partial sums commute): every GEMM that produces a residual-stream snapshot accumulates the rows' sums of squares in
    // its epilogue and every consumer turns them into rstd on the fly — no separate pass re-reads the 63 MB snapshots.
    // ... which holds while the GEMMs that write D-wide rows use 256-wide tiles: small batches switch them to 128-wide tiles (four
    // partial sums per row at D = 512, whose atomic order would make the rollout differ from run to run in the last bit) and keep the
    // separate row passes instead.
    const int fss = (c->fuse_ss && d4_prec_split(c->cfg.precision) && M > 128 && D <= 512 && (D % 4) == 0 && d4_gemm_pair_default() &&
                     (D <= 256 || d4_gemm_pair_bn(M, D) == 256)) ? 1 : 0;
    auto xss = [&](int j) { return c->b.x_rstd + (long long)j * M; };
    if (fss) {
        D4_CUDA_OK(cudaMemsetAsync(hrs(1), 0, (size_t)(c->n_hid - 1) * M * 4, s));
        D4_CUDA_OK(cudaMemsetAsync(c->b.x_rstd, 0, (size_t)L * M * 4, s));
        D4_TRY(d4_row_sumsq(hid(0), D, M, D, hrs(0), s));
    } else {
        D4_TRY(d4_row_rstd(hid(0), D, rowmap_identity(), M, D, hrs(0), s));
    }
    {   // value residual (reference dreamer4.py:3026-3027)
        GemmArgs g = gemm_args(hid(0), D, nullptr, D, c->b.v0, Dkv, M, Dkv, D);
        g.row_scale = hrs(0); g.rs_mode = fss;
        D4_TRY(d4_engine_gemm(c, g, c->vr_w, 0, s));
    }

    // ---- layers (reference dreamer4.py:3043-3223)
    // A denoise pass returns the latent prediction only, which reads the nsp spatial tokens of the last hidden state.  Time attention,
    // feed-forward and the attention-residual pools are per token; only space attention mixes the tokens of a frame.  So everything
    // AFTER the attention of the last space layer `ls` is needed for the spatial rows R only: that layer's out-projection, feed-forward
    // and pool, and every (time) layer after it completely - its fused projection, its time attention (K1 over the spatial tokens'
    // streams only: 4/15 of the cache read) - run on B * nsp of the B * S rows, in place in the (B, S, .) buffers through row maps.  The
    // snapshots of those steps are only valid at R, which is all the later pools read.  (The clean pass needs every row: it appends
    // every token's keys / values and the agent token cross-attends to all of them.)
    int ls = -1;
    for (int i = 0; i < L; ++i) if (!c->is_time[i]) ls = i;
    const bool cone = c->trim_final && c->trim_cone && pred_out && !agent_out && !commit && ls >= 0 && (32 % nsp) == 0 && c->cfg.time_attn_variant == 1;
    const RowMap R = rowmap(nsp, S, 1);
    const int Mr = B * nsp;
    float* rs_c = c->b.fin_rstd2;                       // compact rstd of the current restricted GEMM input
    auto stat_R = [&](const float* x, float* full) { return d4_row_stat_map(x, D, R, Mr, D, rs_c, full, fss, s); };
    const float* x_in = hid(0); const float* x_in_rstd = hrs(0);
    int ti = 0;
    for (int i = 0; i < L; ++i) {
        const bool whole = cone && i > ls;              // the whole layer on R
        const bool tail = cone && i >= ls;              // out-projection, feed-forward and pool on R
        {
            GemmArgs g = gemm_args(x_in, D, nullptr, D, c->b.qkvgm, c->ldq, whole ? Mr : M, c->NQ, D);
            g.row_scale = whole ? rs_c : x_in_rstd; g.rs_mode = whole ? 0 : fss; g.bias = c->attn[i].b;
            if (whole) { g.amap = R; g.cmap = R; }
            D4_TRY(d4_engine_gemm(c, g, c->attn[i].w, 0, s));
        }
        const int off_k = Dq, off_v = Dq + Dkv, off_g = Dq + 2 * Dkv, off_m = Dq + 2 * Dkv + hq;
        if (c->is_time[i]) {
            TimeAttnArgs a; memset(&a, 0, sizeof(a));
            a.M = whole ? Mr : M; a.hkv = h; a.g = hq / h; a.d = d; a.t = t; a.Tmax = c->cfg.max_time;
            if (whole) a.tmap = R;
            a.qkvgm = c->b.qkvgm; a.ld = c->ldq; a.off_k = off_k; a.off_v = off_v; a.off_g = off_g; a.off_m = off_m;
            a.v0 = c->b.v0; a.ldv0 = Dkv; a.k_gamma = c->attn[i].k_gamma; a.inv_freq = c->inv_freq;
            const long long per = (long long)c->cfg.max_batch * S * h * c->cfg.max_time * d;
            a.kcache = c->kv + (long long)(ti * 2 + 0) * per; a.vcache = c->kv + (long long)(ti * 2 + 1) * per;
            a.out = c->b.attn_o; a.ldo = Dq; a.scale = att_scale; a.softclamp = c->cfg.softclamp; a.commit = commit;
            a.variant = c->cfg.time_attn_variant;
            const double kbytes = (double)a.M * h * d * 4.0 * (2.0 * t + 4.0 + (commit ? 2.0 : 0.0));
            const int ph = d4_prof_begin(c, D4_CLS_TIME_ATTN, kbytes, s);
            const int rc = d4_time_attn(a, s);
            d4_prof_end(c, ph, s);
            D4_TRY(rc);
            ++ti;
        } else {
            SmallAttnArgs a; memset(&a, 0, sizeof(a));
            a.nb = B; a.hkv = h; a.g = hq / h; a.d = d; a.nq = S; a.n = S;
            a.q = c->b.qkvgm; a.q_sb = (long long)S * c->ldq; a.q_si = c->ldq;
            a.k = c->b.qkvgm + off_k; a.k_sb = a.q_sb; a.k_sj = c->ldq;
            a.v = c->b.qkvgm + off_v; a.v_sb = a.q_sb; a.v_sj = c->ldq;
            a.k_gamma = c->attn[i].k_gamma;
            a.v0 = c->b.v0; a.v0_sb = (long long)S * Dkv; a.v0_sj = Dkv;
            a.mix = c->b.qkvgm + off_m; a.mix_sb = a.q_sb; a.mix_sj = c->ldq;
            a.gate = c->b.qkvgm + off_g; a.gate_sb = a.q_sb; a.gate_si = c->ldq;
            a.out = c->b.attn_o; a.out_sb = (long long)S * Dq; a.out_si = Dq;
            a.scale = att_scale; a.softclamp = c->cfg.softclamp; a.mask_agent = 1; a.belief = 1;
            a.allow_tensor = (c->cfg.precision != D4_PREC_FP32) && c->space_mma;
            { const int ph = d4_prof_begin(c, D4_CLS_SMALL_ATTN, 0.0, s); const int rc = d4_small_attn(a, s); d4_prof_end(c, ph, s); D4_TRY(rc); }
        }
        if (tail) {
            {
                GemmArgs g = gemm_args(c->b.attn_o, Dq, nullptr, Dq, hid(2 * i + 1), D, Mr, D, Dq);
                g.amap = R; g.cmap = R; g.residual = x_in; g.ldr = D;
                D4_TRY(d4_engine_gemm(c, g, c->attn[i].w_out, 0, s));
            }
            D4_TRY(stat_R(hid(2 * i + 1), hrs(2 * i + 1)));
            D4_TRY(run_ff(c, c->ff[i], Mr, hid(2 * i + 1), D, R, rs_c, 0, hid(2 * i + 2), D, R, nullptr, s));
            D4_TRY(stat_R(hid(2 * i + 2), hrs(2 * i + 2)));
            if (i != L - 1) {
                D4_TRY(run_pool_rows(c, c->pools[i], B, nsp, 1, hid(2 * i + 2), 1, fss, 2 * i + 3, c->b.x_cur, s));
                D4_TRY(stat_R(c->b.x_cur, nullptr));
                x_in = c->b.x_cur; x_in_rstd = rs_c;
            }
            continue;
        }
        {
            GemmArgs g = gemm_args(c->b.attn_o, Dq, nullptr, Dq, hid(2 * i + 1), D, M, D, Dq);
            g.residual = x_in; g.ldr = D; g.ss_out = fss ? hrs(2 * i + 1) : nullptr;
            D4_TRY(d4_engine_gemm(c, g, c->attn[i].w_out, 0, s));
        }
        if (!fss) D4_TRY(d4_row_rstd(hid(2 * i + 1), D, rowmap_identity(), M, D, hrs(2 * i + 1), s));
        D4_TRY(run_ff(c, c->ff[i], M, hid(2 * i + 1), D, rowmap_identity(), hrs(2 * i + 1), fss, hid(2 * i + 2), D, rowmap_identity(),
                      fss ? hrs(2 * i + 2) : nullptr, s));
        if (!fss) D4_TRY(d4_row_rstd(hid(2 * i + 2), D, rowmap_identity(), M, D, hrs(2 * i + 2), s));
        if (i != L - 1) {
            D4_TRY(run_pool(c, c->pools[i], M, hid(2 * i + 2), hrs(2 * i + 2), fss, fss, 2 * i + 3, c->b.x_cur, fss ? xss(i) : nullptr, s));
            if (!fss) D4_TRY(d4_row_rstd(c->b.x_cur, D, rowmap_identity(), M, D, xss(i), s));
            x_in = c->b.x_cur; x_in_rstd = xss(i);
        }
    }

    // ---- what follows the layers only matters for the rows a pass's outputs read: the agent token's cross attention + feed-forward
    // and its row of the final pool on a pass that returns the agent embedding (the clean pass), the spatial tokens' rows of the
    // final pool on a pass that returns the latent prediction (a denoise pass).  The final pool is per token and its K / V projection
    // over all 2L+1 snapshots is the largest single GEMM of the pass (17 of the 80 pool units at depth 8): 4/15 resp. 1/15 of it.
    const bool trim = c->trim_final && ((agent_out != nullptr) != (pred_out != nullptr));
    if (trim && pred_out) {
        float* xs = c->b.fin_x;
        D4_TRY(d4_gather_rows(hid(2 * L), D, rowmap(nsp, S, 1), (long long)B * nsp, D, xs, D, s));
        D4_TRY(run_pool_rows(c, c->pool_final, B, nsp, 1, xs, 0, fss, c->n_hid, xs, s));
        D4_TRY(d4_rmsnorm_rows(xs, D, rowmap_identity(), c->lp_norm0, B * nsp, D, c->b.sp_n, D, s));
        return finish_latent_pred(c, B, pred_out, s);
    }
    if (trim && agent_out) {
        float* xa = c->b.fin_x;
        D4_TRY(d4_gather_rows(hid(2 * L), D, rowmap(1, S, S - 1), B, D, xa, D, s));
        D4_TRY(d4_row_rstd(xa, D, rowmap_identity(), B, D, c->b.ag_rstd, s));
        {
            GemmArgs g = gemm_args(xa, D, nullptr, D, c->b.fa_q, c->ldfa, B, Dq + hq, D);
            g.row_scale = c->b.ag_rstd;
            D4_TRY(d4_engine_gemm(c, g, c->fa.w_qg, 0, s));
            GemmArgs g2 = gemm_args(hid(2 * L), D, nullptr, D, c->b.fa_kv, 2 * Dkv, M, 2 * Dkv, D);
            g2.row_scale = hrs(2 * L); g2.rs_mode = fss;
            D4_TRY(d4_engine_gemm(c, g2, c->fa.w_kv, 0, s));
            SmallAttnArgs a; memset(&a, 0, sizeof(a));
            a.nb = B; a.hkv = h; a.g = hq / h; a.d = d; a.nq = 1; a.n = S - 1;
            a.q = c->b.fa_q; a.q_sb = c->ldfa; a.q_si = 0;
            a.k = c->b.fa_kv; a.k_sb = (long long)S * 2 * Dkv; a.k_sj = 2 * Dkv;
            a.v = c->b.fa_kv + Dkv; a.v_sb = a.k_sb; a.v_sj = a.k_sj;
            a.k_gamma = c->fa.k_gamma;
            a.gate = c->b.fa_q + Dq; a.gate_sb = c->ldfa; a.gate_si = 0;
            a.out = c->b.fa_att; a.out_sb = Dq; a.out_si = 0;
            a.scale = att_scale;
            { const int ph = d4_prof_begin(c, D4_CLS_SMALL_ATTN, 0.0, s); const int rc = d4_small_attn(a, s); d4_prof_end(c, ph, s); D4_TRY(rc); }
            GemmArgs g3 = gemm_args(c->b.fa_att, Dq, nullptr, Dq, xa, D, B, D, Dq);
            g3.residual = xa; g3.ldr = D;
            D4_TRY(d4_engine_gemm(c, g3, c->fa.w_out, 0, s));
        }
        D4_TRY(d4_row_rstd(xa, D, rowmap_identity(), B, D, c->b.ag_rstd, s));
        D4_TRY(run_ff(c, c->fa_ff, B, xa, D, rowmap_identity(), c->b.ag_rstd, 0, xa, D, rowmap_identity(), nullptr, s));
        return run_pool_rows(c, c->pool_final, B, 1, S - 1, xa, 0, fss, c->n_hid, agent_out, s);
    }

    // ---- final agent-token cross attention + feed-forward (reference dreamer4.py:3227-3238)
    float* xf = c->b.x_cur;
    D4_CUDA_OK(cudaMemcpyAsync(xf, hid(2 * L), MD * 4, cudaMemcpyDeviceToDevice, s));
    const RowMap agent_rows = rowmap(1, S, S - 1);
    D4_TRY(d4_row_rstd(xf, D, agent_rows, B, D, c->b.ag_rstd, s));
    {
        GemmArgs g = gemm_args(xf, D, nullptr, D, c->b.fa_q, c->ldfa, B, Dq + hq, D);
        g.amap = agent_rows; g.row_scale = c->b.ag_rstd;
        D4_TRY(d4_engine_gemm(c, g, c->fa.w_qg, 0, s));
        GemmArgs g2 = gemm_args(hid(2 * L), D, nullptr, D, c->b.fa_kv, 2 * Dkv, M, 2 * Dkv, D);
        g2.row_scale = hrs(2 * L); g2.rs_mode = fss;
        D4_TRY(d4_engine_gemm(c, g2, c->fa.w_kv, 0, s));
        SmallAttnArgs a; memset(&a, 0, sizeof(a));
        a.nb = B; a.hkv = h; a.g = hq / h; a.d = d; a.nq = 1; a.n = S - 1;
        a.q = c->b.fa_q; a.q_sb = c->ldfa; a.q_si = 0;
        a.k = c->b.fa_kv; a.k_sb = (long long)S * 2 * Dkv; a.k_sj = 2 * Dkv;
        a.v = c->b.fa_kv + Dkv; a.v_sb = a.k_sb; a.v_sj = a.k_sj;
        a.k_gamma = c->fa.k_gamma;
        a.gate = c->b.fa_q + Dq; a.gate_sb = c->ldfa; a.gate_si = 0;
        a.out = c->b.fa_att; a.out_sb = Dq; a.out_si = 0;
        a.scale = att_scale;
        { const int ph = d4_prof_begin(c, D4_CLS_SMALL_ATTN, 0.0, s); const int rc = d4_small_attn(a, s); d4_prof_end(c, ph, s); D4_TRY(rc); }
        GemmArgs g3 = gemm_args(c->b.fa_att, Dq, nullptr, Dq, xf, D, B, D, Dq);
        g3.residual = xf; g3.ldr = D; g3.cmap = agent_rows;
        D4_TRY(d4_engine_gemm(c, g3, c->fa.w_out, 0, s));
    }
    D4_TRY(d4_row_rstd(xf, D, agent_rows, B, D, c->b.ag_rstd, s));
    D4_TRY(run_ff(c, c->fa_ff, B, xf, D, agent_rows, c->b.ag_rstd, 0, xf, D, agent_rows, nullptr, s));
    // final attention-residual pool over all 2L+1 hiddens (reference dreamer4.py:3242-3243)
    D4_TRY(d4_row_rstd(xf, D, rowmap_identity(), M, D, xss(L), s));          // xf = last snapshot with the agent rows updated: own pass
    D4_TRY(run_pool(c, c->pool_final, M, xf, xss(L), 0, fss, c->n_hid, xf, nullptr, s));

    // ---- outputs: agent embedding and the latent prediction (reference dreamer4.py:7251, 4830-4834)
    if (agent_out) D4_TRY(d4_copy_rows(xf + (long long)(S - 1) * D, (long long)S * D, agent_out, D, B, D, s));
    if (pred_out) {
        D4_TRY(d4_rmsnorm_rows(xf, D, rowmap(nsp, S, 1), c->lp_norm0, B * nsp, D, c->b.sp_n, D, s));
        return finish_latent_pred(c, B, pred_out, s);
    }
    return 0;
}

int check_ready(d4_ctx* c, int B, int t) {
    if (!c) return d4_fail("null ctx");
    if (!c->bound) return d4_fail("weights not bound: call d4_bind() after d4_set_weight()");
    if (!c->ws) return d4_fail("buffers not set: call d4_set_buffers()");
    if (B < 1 || B > c->cfg.max_batch) return d4_fail("batch %d outside 1..max_batch=%d", B, c->cfg.max_batch);
    if (t < 0 || t >= c->cfg.max_time) return d4_fail("frame index %d outside the KV capacity max_time=%d", t, c->cfg.max_time);
    return 0;
}

}  // namespace

extern "C" int d4_pass(d4_ctx* c, int B, const float* latent, int signal_level, int step_size_log2, const int64_t* prev_actions,
                       int64_t pa_stride, const int64_t* tasks, int t, int commit_kv, float* pred_out, float* agent_out, void* stream) {
    D4_TRY(check_ready(c, B, t));
    return run_pass(c, B, latent, signal_level, step_size_log2, prev_actions, pa_stride, tasks, t, commit_kv, pred_out, agent_out,
                    static_cast<cudaStream_t>(stream));
}

// d4_pass with per-dream signal levels / step sizes (DynamicsWorldModel.forward's (b, t) signal_levels and (b) step_sizes, reference
// dreamer4.py:6912-6942): one frame of the inference branch
extern "C" int d4_pass_ex(d4_ctx* c, int B, const float* latent, const int64_t* signal_levels, const int64_t* step_sizes_log2,
                          const int64_t* prev_actions, int64_t pa_stride, const int64_t* tasks, int t, int commit_kv, float* pred_out,
                          float* agent_out, void* stream) {
    D4_TRY(check_ready(c, B, t));
    if (!signal_levels || !step_sizes_log2) return d4_fail("d4_pass_ex: signal_levels and step_sizes_log2 (B) are required");
    return run_pass(c, B, latent, 0, 0, prev_actions, pa_stride, tasks, t, commit_kv, pred_out, agent_out, static_cast<cudaStream_t>(stream),
                    signal_levels, step_sizes_log2);
}

// A head MLP on caller rows (the modules the reference exposes as model.policy_head / model.value_head / the terminal head:
// x-mlps create_mlp, reference dreamer4.py:4950-4956, 5083-5101): x (M, dim_in) -> out (M, dim_out), M <= max_batch rows per call.
extern "C" int d4_head_forward(d4_ctx* c, int which, const float* x, int M, float* out, void* stream) {
    if (!c || !x || !out) return d4_fail("d4_head_forward: null argument");
    if (!c->bound || !c->ws) return d4_fail("d4_head_forward: context not ready (d4_bind / d4_set_buffers)");
    const MlpW* m = which == 0 ? &c->policy : which == 1 ? &c->value : which == 2 ? &c->terminal : nullptr;
    if (!m || m->layers == 0) return d4_fail("d4_head_forward: head %d is not part of this model", which);
    if (M < 1 || M > c->cfg.max_batch) return d4_fail("d4_head_forward: %d rows outside 1..max_batch=%d", M, c->cfg.max_batch);
    return d4_mlp_forward(c, *m, x, m->dims[0], M, c->b.hbuf0, c->b.hbuf1, out, m->dims[m->layers], which != 2, static_cast<cudaStream_t>(stream));
}

// One frame at cache position t: passes first_step..num_steps of the denoising schedule (the last one is the clean pass that
// commits the frame's keys/values), then the heads.  first_step = 0 is d4_frame; first_step = num_steps is d4_observe, where
// io->noise_latent already holds the clean latent and only that last pass runs.
static int frame_body(d4_ctx* c, int B, int t, int num_steps, float discrete_temperature, const d4_frame_io* io, void* stream, int first_step) {
    D4_TRY(check_ready(c, B, t));
    if (!io || !io->noise_latent || !io->latents) return d4_fail("d4_frame: noise_latent and latents are required");
    if (num_steps < 1 || num_steps > c->cfg.max_steps || (num_steps & (num_steps - 1)) || c->cfg.max_steps % num_steps)
        return d4_fail("d4_frame: num_steps=%d must be a power of two dividing max_steps=%d", num_steps, c->cfg.max_steps);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int D = c->D, N = c->N, Dl = c->Dl;
    const long long nlat = (long long)B * N * Dl;
    const int step_size = c->cfg.max_steps / num_steps;
    int step_log2 = 0; while ((1 << step_log2) < step_size) ++step_log2;
    float* x = c->b.lat_x;
    D4_CUDA_OK(cudaMemcpyAsync(x, io->noise_latent, nlat * 4, cudaMemcpyDeviceToDevice, s));
    // denoising passes + the clean pass that commits this frame's keys/values (reference dreamer4.py:6484-6580)
    for (int step = first_step; step <= num_steps; ++step) {
        const bool last = (step == num_steps);
        const int signal = std::min(step * step_size, c->cfg.max_steps - 1);
        D4_TRY(run_pass(c, B, x, signal, step_log2, io->prev_actions, io->pa_stride, io->tasks, t, last ? 1 : 0,
                        last ? nullptr : c->b.pred, last ? c->b.agent : nullptr, s));
        if (!last) {
            const float tau = (float)signal / (float)c->cfg.max_steps;
            D4_TRY(d4_flow_step(x, c->b.pred, nlat, 1.f - tau, (float)step_size / (float)c->cfg.max_steps, s));
        }
    }
    D4_TRY(d4_store_latents(x, io->latents, B, (long long)N * Dl, io->latents_bs, s));
    if (io->agent_embed) D4_TRY(d4_copy_rows(c->b.agent, D, io->agent_embed, io->agent_bs, B, D, s));
    // reward head: Ensemble member 0 = RMSNorm -> Linear(no bias) -> HL-Gauss expectation (reference dreamer4.py:6598-6601)
    if (io->rewards) {
        D4_TRY(d4_row_rstd(c->b.agent, D, rowmap_identity(), B, D, c->b.ag_rstd, s));
        GemmArgs g = gemm_args(c->b.agent, D, c->reward_w, D, c->b.bins, c->cfg.reward_bins, B, c->cfg.reward_bins, D);
        g.row_scale = c->b.ag_rstd;
        D4_TRY(d4_engine_gemm(c, g, plain(c->reward_w), 1, s));
        D4_TRY(d4_hl_gauss_decode(c->b.bins, c->cfg.reward_bins, B, c->cfg.reward_bins, c->reward_centers, io->rewards, io->rewards_bs, s));
    }
    // terminal head (reference dreamer4.py:6605-6616)
    if (c->cfg.predict_terminals && io->terminal_uniform && io->lens && io->terminals) {
        D4_TRY(d4_mean_tokens(x, B, N, Dl, c->b.term_in, s));
        D4_TRY(d4_mlp_forward(c, c->terminal, c->b.term_in, Dl, B, c->b.hbuf0, c->b.hbuf1, c->b.bins, 1, 0, s));
        D4_TRY(d4_terminal_update(c->b.bins, 1, io->terminal_uniform, B, t, reinterpret_cast<long long*>(io->lens), io->terminals, s));
    }
    // policy head -> logits -> gumbel-argmax; value head (reference dreamer4.py:6628-6662)
    if (c->has_actions && io->actions) {
        if (!io->action_uniform || !io->log_probs) return d4_fail("d4_frame: action_uniform and log_probs are required with actions");
        float* pe = (c->policy.layers & 1) ? c->b.hbuf0 : c->b.hbuf1;   // buffer not used by the last hidden layer
        // In the tf32x3 engine mode the policy MLP runs on the 3xTF32 tensor-core path like the value head.  Measured on the
        // BASELINE architectures (tests/test_gpu_parity.py::test_baseline_architectures_match_oracle): the logit error against
        // the fp32 oracle is the same (1e-4 .. 2.4e-4 at logits of +-10) whether this MLP runs exact-fp32 FMA or 3xTF32 — it
        // is set by the agent embedding coming out of the transformer — and sampled action indices are bit-identical in both.
        // The final unembedding (N = number of actions) is always exact-fp32 FMA.
        D4_TRY(d4_mlp_forward(c, c->policy, c->b.agent, D, B, c->b.hbuf0, c->b.hbuf1, pe, c->cfg.policy_hidden, 1, s));
        GemmArgs g = gemm_args(pe, c->cfg.policy_hidden, c->unembed, c->unembed_ld, c->b.logits, c->ldlog, B, c->A_total, c->cfg.policy_hidden);
        D4_TRY(d4_engine_gemm(c, g, plain(c->unembed), 1, s));
        if (io->logits) D4_TRY(d4_copy_rows(c->b.logits, c->ldlog, io->logits, io->logits_bs, B, c->A_total, s));
        const float inv_temp = 1.f / fmaxf(discrete_temperature, 1e-10f);
        D4_TRY(d4_sample_actions(c->b.logits, c->ldlog, io->action_uniform, c->A_total, B, c->na, c->b.sizes_offs, inv_temp,
                                 reinterpret_cast<long long*>(io->actions), io->actions_bs, io->log_probs, io->log_probs_bs, s));
        if (io->values) {
            D4_TRY(d4_mlp_forward(c, c->value, c->b.agent, D, B, c->b.hbuf0, c->b.hbuf1, c->b.bins, c->cfg.value_bins, 1, s));
            D4_TRY(d4_hl_gauss_decode(c->b.bins, c->cfg.value_bins, B, c->cfg.value_bins, c->value_centers, io->values, io->values_bs, s));
        }
    }
    return 0;
}

// ---- CUDA-graph replay of a frame (see d4_ctx::use_graphs)
namespace {
// rows of `width` bytes between a dense staging buffer and the caller's strided rows
inline cudaError_t rows_copy(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, int rows, cudaStream_t s) {
    if (dpitch == width && spitch == width) return cudaMemcpyAsync(dst, src, width * rows, cudaMemcpyDeviceToDevice, s);
    return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, cudaMemcpyDeviceToDevice, s);
}
}  // namespace

static int frame_impl(d4_ctx* c, int B, int t, int num_steps, float discrete_temperature, const d4_frame_io* io, void* stream, int first_step) {
    if (!c || !c->use_graphs || c->prof_on || !io || (long long)B * c->S > c->graph_max_rows || c->graphs.size() >= 4096)
        return frame_body(c, B, t, num_steps, discrete_temperature, io, stream, first_step);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool term = c->cfg.predict_terminals && io->terminal_uniform && io->lens && io->terminals;
    const bool act = c->has_actions && io->actions;
    long long flags = 0;
    const void* opt[] = {io->prev_actions, io->tasks, io->agent_embed, io->rewards, io->values, io->logits, io->action_uniform, io->log_probs};
    for (int i = 0; i < 8; ++i) flags |= (long long)(opt[i] != nullptr) << i;
    flags |= (long long)term << 8 | (long long)act << 9;
    uint32_t tbits; memcpy(&tbits, &discrete_temperature, 4);
    const std::array<long long, 6> key = {first_step == 0 ? 0 : 1, B, t, num_steps, (long long)tbits, flags};
    auto& fg = c->graphs[key];
    if (!fg.seen || fg.direct) {           // first use of this key (or a stream that cannot be captured): run directly
        fg.seen = true;
        return frame_body(c, B, t, num_steps, discrete_temperature, io, stream, first_step);
    }
    const auto& g = c->gio;
    const long long nl = (long long)c->N * c->Dl, A = c->A_total, na = c->na;
    d4_frame_io st; memset(&st, 0, sizeof(st));
    st.noise_latent = g.noise; st.latents = g.latents; st.latents_bs = nl;
    if (io->action_uniform) st.action_uniform = g.act_u;
    if (io->terminal_uniform) st.terminal_uniform = g.term_u;
    if (io->prev_actions) { st.prev_actions = reinterpret_cast<const int64_t*>(g.prev_actions); st.pa_stride = na; }
    if (io->tasks) st.tasks = reinterpret_cast<const int64_t*>(g.tasks);
    if (io->agent_embed) { st.agent_embed = g.agent; st.agent_bs = c->D; }
    if (io->rewards) { st.rewards = g.rewards; st.rewards_bs = 1; }
    if (io->values) { st.values = g.values; st.values_bs = 1; }
    if (io->actions) { st.actions = reinterpret_cast<int64_t*>(g.actions); st.actions_bs = na; }
    if (io->log_probs) { st.log_probs = g.logp; st.log_probs_bs = na; }
    if (io->logits) { st.logits = g.logits; st.logits_bs = A; }
    if (term) { st.lens = reinterpret_cast<int64_t*>(g.lens); st.terminals = g.terminals; }
    D4_TRY(graph_stream(c));
    cudaStream_t gs = c->gstream;
    if (!fg.exec) {           // second use: capture the frame on the staging rows (on the engine's own stream), instantiate
        if (!c->bound || !c->ws) return frame_body(c, B, t, num_steps, discrete_temperature, io, stream, first_step);   // reports the error
        const long long l0 = d4_launches_;
        if (cudaStreamBeginCapture(gs, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();
            fg.direct = true;
            return frame_body(c, B, t, num_steps, discrete_temperature, io, stream, first_step);
        }
        const int rc = frame_body(c, B, t, num_steps, discrete_temperature, &st, gs, first_step);
        cudaGraph_t graph = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(gs, &graph);
        const long long captured = d4_launches_ - l0;
        d4_launches_ = l0;                                   // nothing ran yet
        if (rc != 0) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); return rc; }
        if (ce != cudaSuccess || !graph) { cudaGetLastError(); return d4_fail("d4_frame: stream capture failed: %s", cudaGetErrorString(ce)); }
        const cudaError_t ie = cudaGraphInstantiate(&fg.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) { fg.exec = nullptr; cudaGetLastError(); return d4_fail("d4_frame: cudaGraphInstantiate failed: %s", cudaGetErrorString(ie)); }
        fg.launches = captured;
    }
    // copy in (caller's stream) -> replay (engine's stream) -> copy out (caller's stream)
    D4_CUDA_OK(cudaMemcpyAsync(g.noise, io->noise_latent, (size_t)B * nl * 4, cudaMemcpyDeviceToDevice, s));
    if (io->action_uniform) D4_CUDA_OK(cudaMemcpyAsync(g.act_u, io->action_uniform, (size_t)B * A * 4, cudaMemcpyDeviceToDevice, s));
    if (io->terminal_uniform) D4_CUDA_OK(cudaMemcpyAsync(g.term_u, io->terminal_uniform, (size_t)B * 4, cudaMemcpyDeviceToDevice, s));
    if (io->prev_actions) D4_CUDA_OK(rows_copy(g.prev_actions, na * 8, io->prev_actions, (size_t)io->pa_stride * 8, na * 8, B, s));
    if (io->tasks) D4_CUDA_OK(cudaMemcpyAsync(g.tasks, io->tasks, (size_t)B * 8, cudaMemcpyDeviceToDevice, s));
    if (term) {
        D4_CUDA_OK(cudaMemcpyAsync(g.lens, io->lens, (size_t)B * 8, cudaMemcpyDeviceToDevice, s));
        D4_CUDA_OK(cudaMemcpyAsync(g.terminals, io->terminals, (size_t)B, cudaMemcpyDeviceToDevice, s));
    }
    // the graph runs on the engine's stream between two events: after everything the caller has queued (inputs, the previous frame),
    // before anything the caller queues next
    D4_CUDA_OK(cudaEventRecord(c->gev_in, s));
    D4_CUDA_OK(cudaStreamWaitEvent(gs, c->gev_in, 0));
    D4_CUDA_OK(cudaGraphLaunch(fg.exec, gs));
    D4_CUDA_OK(cudaEventRecord(c->gev_out, gs));
    D4_CUDA_OK(cudaStreamWaitEvent(s, c->gev_out, 0));
    d4_launches_ += fg.launches; ++c->graph_replays;
    D4_CUDA_OK(rows_copy(io->latents, (size_t)io->latents_bs * 4, g.latents, nl * 4, nl * 4, B, s));
    if (io->agent_embed) D4_CUDA_OK(rows_copy(io->agent_embed, (size_t)io->agent_bs * 4, g.agent, (size_t)c->D * 4, (size_t)c->D * 4, B, s));
    if (io->rewards) D4_CUDA_OK(rows_copy(io->rewards, (size_t)io->rewards_bs * 4, g.rewards, 4, 4, B, s));
    if (act) {
        D4_CUDA_OK(rows_copy(io->actions, (size_t)io->actions_bs * 8, g.actions, na * 8, na * 8, B, s));
        if (io->log_probs) D4_CUDA_OK(rows_copy(io->log_probs, (size_t)io->log_probs_bs * 4, g.logp, na * 4, na * 4, B, s));
        if (io->logits) D4_CUDA_OK(rows_copy(io->logits, (size_t)io->logits_bs * 4, g.logits, A * 4, A * 4, B, s));
        if (io->values) D4_CUDA_OK(rows_copy(io->values, (size_t)io->values_bs * 4, g.values, 4, 4, B, s));
    }
    if (term) {
        D4_CUDA_OK(cudaMemcpyAsync(io->lens, g.lens, (size_t)B * 8, cudaMemcpyDeviceToDevice, s));
        D4_CUDA_OK(cudaMemcpyAsync(io->terminals, g.terminals, (size_t)B, cudaMemcpyDeviceToDevice, s));
    }
    return 0;
}

extern "C" int d4_frame(d4_ctx* c, int B, int t, int num_steps, float discrete_temperature, const d4_frame_io* io, void* stream) {
    return frame_impl(c, B, t, num_steps, discrete_temperature, io, stream, 0);
}

extern "C" int d4_observe(d4_ctx* c, int B, int t, int num_steps, float discrete_temperature, const d4_frame_io* io, void* stream) {
    return frame_impl(c, B, t, num_steps, discrete_temperature, io, stream, num_steps);
}

// ------------------------------------------------------------------------------------------------ stand-alone operators
extern "C" int d4_time_attn_decode(int M, int heads, int query_heads, int dim_head, int t, int Tmax, const float* qkvgm, int64_t ld,
                                   const float* v0, const float* k_gamma, const float* inv_freq, float* kcache, float* vcache, float* out,
                                   float softclamp, int commit, int variant, void* stream) {
    if (query_heads % heads) return d4_fail("query_heads must be a multiple of heads");
    TimeAttnArgs a; memset(&a, 0, sizeof(a));
    const int Dq = query_heads * dim_head, Dkv = heads * dim_head;
    a.M = M; a.hkv = heads; a.g = query_heads / heads; a.d = dim_head; a.t = t; a.Tmax = Tmax;
    a.qkvgm = qkvgm; a.ld = ld; a.off_k = Dq; a.off_v = Dq + Dkv; a.off_g = Dq + 2 * Dkv; a.off_m = Dq + 2 * Dkv + query_heads;
    a.v0 = v0; a.ldv0 = Dkv; a.k_gamma = k_gamma; a.inv_freq = inv_freq; a.kcache = kcache; a.vcache = vcache;
    a.out = out; a.ldo = Dq; a.scale = 1.f / sqrtf((float)dim_head); a.softclamp = softclamp; a.commit = commit; a.variant = variant;
    return d4_time_attn(a, static_cast<cudaStream_t>(stream));
}

extern "C" int d4_linear(int precision, int M, int N, int K, const float* A, int64_t lda, const float* W, int64_t ldw, const float* W_lo,
                         const float* bias, const float* row_scale, const float* residual, int64_t ldr, int act, float* C, int64_t ldc,
                         void* stream) {
    GemmArgs g = gemm_args(A, lda, W, ldw, C, ldc, M, N, K);
    g.bias = bias; g.row_scale = row_scale; g.residual = residual; g.ldr = ldr; g.act = act; g.W_lo = W_lo;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (precision == D4_PREC_FP32) return d4_gemm_skinny_supported(g) ? d4_gemm_skinny(g, s) : d4_gemm_simt(g, s);      // <= 32 rows: the weight-streaming kernel
    if (precision == D4_PREC_F16X3) {          // experimental: W / W_lo point to fp16 (N, ldw) hi / lo words of weights pre-scaled to rms ~ 1
        if (!W_lo) return d4_fail("d4_linear: f16x3 needs W_lo");
        return d4_gemm_f16x3(g, 1.f, 0, s);    // the caller folds 1 / q into row_scale
    }
    if (!d4_gemm_tc_supported(g)) return d4_fail("d4_linear: shape/alignment not supported by the tcgen05 path (need K%%4==0, 16B-aligned rows)");
    if (precision == D4_PREC_TF32X3) { if (!W_lo) return d4_fail("d4_linear: tf32x3 needs W_lo"); return d4_gemm_tc(g, 3, s); }
    return d4_gemm_tc(g, 1, s);
}

// ================================================================================================ generic transformer context
// The video tokenizer's encoder and decoder are AxialSpaceTimeTransformers like the dynamics model's (reference
// dreamer4.py:3908-3929, 3595-3607) with other token counts: S = patches + latents per frame, the last num_special of them
// special (the encoder's latent tokens: patches cannot attend to them, they cross-attend to the patches once more at the end,
// 3912-3918 / 3227-3238; the decoder keeps the library default of one), and a final RMSNorm (2774, 3247).  d4_tf_step is the
// layer loop of run_pass on caller-assembled tokens, with the attention inside a frame on frame_attn.cu and the RMS statistics
// as separate row passes.  STATUS: drafted in round 1 after the GPU budget was spent - compiles, not yet run on hardware.

extern "C" int d4_tf_create(const d4_tf_config* cfg, d4_ctx** out) {
    if (!cfg || !out) return d4_fail("d4_tf_create: null argument");
    d4_ctx* c = new d4_ctx();
    memset(&c->cfg, 0, sizeof(c->cfg));
    c->cfg.dim = cfg->dim; c->cfg.depth = cfg->depth; c->cfg.time_block_every = cfg->time_block_every;
    c->cfg.heads = cfg->heads; c->cfg.query_heads = cfg->query_heads; c->cfg.dim_head = cfg->dim_head;
    c->cfg.pool_heads = cfg->pool_heads; c->cfg.pool_dim_head = cfg->pool_dim_head;
    c->cfg.ff_inner = cfg->ff_inner; c->cfg.ff_inner_pad = cfg->ff_inner_pad; c->cfg.ff_act = cfg->ff_act;
    c->cfg.softclamp = cfg->softclamp; c->cfg.max_batch = cfg->max_batch; c->cfg.max_time = cfg->max_time;
    c->cfg.precision = cfg->precision; c->cfg.time_attn_variant = cfg->time_attn_variant;
    c->cfg.max_steps = 1;
    c->tf_mode = true; c->tf_ns = cfg->num_special; c->tf_final_norm = cfg->final_norm;
    c->D = cfg->dim; c->Dl = 0; c->N = 0; c->nsp = 0; c->nreg = 0; c->L = cfg->depth; c->h = cfg->heads; c->hq = cfg->query_heads; c->d = cfg->dim_head;
    c->hp = cfg->pool_heads; c->dp = cfg->pool_dim_head; c->Dp = c->hp * c->dp;
    c->Dq = c->hq * c->d; c->Dkv = c->h * c->d;
    c->inner = cfg->ff_inner; c->inner_pad = cfg->ff_inner_pad;
    c->na = 0; c->has_actions = 0; c->A_total = 0; c->same_len = 0;
    c->S = cfg->tokens_per_frame;
    c->n_hid = 2 * c->L + 1;
    c->y = 0;
    const char* bad = nullptr;
    if (cfg->depth < 1 || cfg->time_block_every < 1) bad = "depth and time_block_every must be positive";
    else if (c->h < 1 || c->hq % c->h != 0) bad = "query_heads must be a multiple of heads";
    else if (c->d % 4 != 0 || c->d > 128) bad = "dim_head must be a multiple of 4, <= 128";
    else if (c->D % 4 != 0) bad = "dim must be a multiple of 4";
    else if (c->S < 2 || c->tf_ns < 1 || c->tf_ns >= c->S) bad = "need 1 <= num_special < tokens_per_frame";
    else if (c->n_hid > 64) bad = "depth > 31 unsupported";
    else if (c->inner_pad < c->inner || c->inner_pad % 4 != 0) bad = "ff_inner_pad must be >= ff_inner and a multiple of 4";
    else if (cfg->max_batch < 1 || cfg->max_time < 1) bad = "max_batch / max_time must be positive";
    if (bad) { delete c; return d4_fail("d4_tf_create: %s", bad); }
    for (int i = 0; i < c->L; ++i) { const int it = ((i + 1) % cfg->time_block_every) == 0; c->is_time.push_back(it); c->y += it; }
    c->NQ = c->Dq + 2 * c->Dkv + c->hq + c->h; c->ldq = round_up(c->NQ, 4);
    c->ldpq = round_up(c->Dp + c->hp, 4);
    c->ldfa = round_up(c->Dq + c->hq, 4);
    c->ldlog = 4;
    c->fuse_pools = false; c->space_mma = false; c->fuse_ss = false; c->use_graphs = false;
    { const char* f = getenv("D4_SKINNY"); c->skinny = f ? atoi(f) != 0 : true; }
    d4_engine_plan(c);
    *out = c;
    return 0;
}

static int tf_bind(d4_ctx* c) {
    Binder B{c};
    const int D = c->D, Dq = c->Dq, Dkv = c->Dkv, h = c->h, hq = c->hq, d = c->d, Dp = c->Dp, hp = c->hp;
    c->vr_w = B.lin("vr.w", (int64_t)Dkv * D);
    c->inv_freq = B.get("inv_freq", d / 2);
    c->tf_final_norm_w = B.get("final_norm", D, !c->tf_final_norm);
    auto bind_ff = [&](const std::string& p, FFW& f) {
        f.w_in = B.lin(p + ".w_in", (int64_t)2 * c->inner * D); f.b_in = B.get(p + ".b_in", 2 * c->inner);
        f.w_out = B.lin(p + ".w_out", (int64_t)D * c->inner_pad); f.b_out = B.get(p + ".b_out", D);
    };
    auto bind_pool = [&](const std::string& p, PoolW& w) {
        w.w_qg = B.lin(p + ".w_qg", (int64_t)(Dp + hp) * D); w.w_kv = B.lin(p + ".w_kv", (int64_t)2 * Dp * D);
        w.k_gamma = B.get(p + ".k_gamma", (int64_t)hp * c->dp); w.w_out = B.lin(p + ".w_out", (int64_t)D * Dp);
    };
    c->attn.assign(c->L, AttnLayerW()); c->ff.assign(c->L, FFW()); c->pools.assign(c->L > 0 ? c->L - 1 : 0, PoolW());
    for (int i = 0; i < c->L; ++i) {
        const std::string p = "L" + std::to_string(i);
        c->attn[i].w = B.lin(p + ".attn.w", (int64_t)c->NQ * D); c->attn[i].b = B.get(p + ".attn.b", c->NQ);
        c->attn[i].k_gamma = B.get(p + ".attn.k_gamma", (int64_t)h * d); c->attn[i].w_out = B.lin(p + ".attn.w_out", (int64_t)D * Dq);
        bind_ff(p + ".ff", c->ff[i]);
        if (i != c->L - 1) bind_pool("P" + std::to_string(i), c->pools[i]);
    }
    bind_pool("PF", c->pool_final);
    c->fa.w_qg = B.lin("FA.w_qg", (int64_t)(Dq + hq) * D); c->fa.w_kv = B.lin("FA.w_kv", (int64_t)2 * Dkv * D);
    c->fa.k_gamma = B.get("FA.k_gamma", (int64_t)h * d); c->fa.w_out = B.lin("FA.w_out", (int64_t)D * Dq);
    bind_ff("FAFF", c->fa_ff);
    if (B.missing) return d4_fail("d4_bind: weight '%s' missing or mis-sized", B.missing);
    c->bound = true;
    return 0;
}

extern "C" int d4_tf_step(d4_ctx* c, int B, const float* tokens_in, int t, float* tokens_out, void* stream) {
    D4_TRY(check_ready(c, B, t));
    if (!c->tf_mode) return d4_fail("d4_tf_step: not a d4_tf_create context");
    if (!tokens_in || !tokens_out) return d4_fail("d4_tf_step: null tokens");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int S = c->S, D = c->D, Dq = c->Dq, Dkv = c->Dkv, h = c->h, hq = c->hq, d = c->d, L = c->L, ns = c->tf_ns;
    const int M = B * S, Ms = B * ns;
    const long long MD = (long long)M * D;
    auto hid = [&](int j) { return c->b.hid + (long long)j * MD; };
    auto hrs = [&](int j) { return c->b.hid_rstd + (long long)j * M; };
    auto xss = [&](int j) { return c->b.x_rstd + (long long)j * M; };
    const float att_scale = 1.f / sqrtf((float)d);
    auto attend = [&](SmallAttnArgs a) {
        a.allow_tensor = (c->cfg.precision != D4_PREC_FP32);          // 3xTF32 mma.sync tiles (frame_attn.cu) in the tensor-core engine modes
        const int ph = d4_prof_begin(c, D4_CLS_SMALL_ATTN, 0.0, s); const int rc = d4_frame_attn(a, s); d4_prof_end(c, ph, s); return rc;
    };

    D4_CUDA_OK(cudaMemcpyAsync(hid(0), tokens_in, MD * 4, cudaMemcpyDeviceToDevice, s));
    D4_TRY(d4_row_rstd(hid(0), D, rowmap_identity(), M, D, hrs(0), s));
    {   // value residual (reference dreamer4.py:3026-3027)
        GemmArgs g = gemm_args(hid(0), D, nullptr, D, c->b.v0, Dkv, M, Dkv, D);
        g.row_scale = hrs(0);
        D4_TRY(d4_engine_gemm(c, g, c->vr_w, 0, s));
    }
    const float* x_in = hid(0); const float* x_in_rstd = hrs(0);
    int ti = 0;
    for (int i = 0; i < L; ++i) {
        {   // fused q | k | v | gate logits | value-residual mix logits of the normed layer input
            GemmArgs g = gemm_args(x_in, D, nullptr, D, c->b.qkvgm, c->ldq, M, c->NQ, D);
            g.row_scale = x_in_rstd; g.bias = c->attn[i].b;
            D4_TRY(d4_engine_gemm(c, g, c->attn[i].w, 0, s));
        }
        const int off_k = Dq, off_v = Dq + Dkv, off_g = Dq + 2 * Dkv, off_m = Dq + 2 * Dkv + hq;
        if (c->is_time[i]) {
            TimeAttnArgs a; memset(&a, 0, sizeof(a));
            a.M = M; a.hkv = h; a.g = hq / h; a.d = d; a.t = t; a.Tmax = c->cfg.max_time;
            a.qkvgm = c->b.qkvgm; a.ld = c->ldq; a.off_k = off_k; a.off_v = off_v; a.off_g = off_g; a.off_m = off_m;
            a.v0 = c->b.v0; a.ldv0 = Dkv; a.k_gamma = c->attn[i].k_gamma; a.inv_freq = c->inv_freq;
            const long long per = (long long)c->cfg.max_batch * S * h * c->cfg.max_time * d;
            a.kcache = c->kv + (long long)(ti * 2 + 0) * per; a.vcache = c->kv + (long long)(ti * 2 + 1) * per;
            a.out = c->b.attn_o; a.ldo = Dq; a.scale = att_scale; a.softclamp = c->cfg.softclamp; a.commit = 1;
            a.variant = c->cfg.time_attn_variant;
            const int ph = d4_prof_begin(c, D4_CLS_TIME_ATTN, (double)M * h * d * 4.0 * (2.0 * t + 6.0), s);
            const int rc = d4_time_attn(a, s);
            d4_prof_end(c, ph, s);
            D4_TRY(rc);
            ++ti;
        } else {
            SmallAttnArgs a; memset(&a, 0, sizeof(a));
            a.nb = B; a.hkv = h; a.g = hq / h; a.d = d; a.nq = S; a.n = S;
            a.q = c->b.qkvgm; a.q_sb = (long long)S * c->ldq; a.q_si = c->ldq;
            a.k = c->b.qkvgm + off_k; a.k_sb = a.q_sb; a.k_sj = c->ldq;
            a.v = c->b.qkvgm + off_v; a.v_sb = a.q_sb; a.v_sj = c->ldq;
            a.k_gamma = c->attn[i].k_gamma;
            a.v0 = c->b.v0; a.v0_sb = (long long)S * Dkv; a.v0_sj = Dkv;
            a.mix = c->b.qkvgm + off_m; a.mix_sb = a.q_sb; a.mix_sj = c->ldq;
            a.gate = c->b.qkvgm + off_g; a.gate_sb = a.q_sb; a.gate_si = c->ldq;
            a.out = c->b.attn_o; a.out_sb = (long long)S * Dq; a.out_si = Dq;
            a.scale = att_scale; a.softclamp = c->cfg.softclamp; a.mask_agent = ns; a.belief = 1;
            D4_TRY(attend(a));
        }
        {
            GemmArgs g = gemm_args(c->b.attn_o, Dq, nullptr, Dq, hid(2 * i + 1), D, M, D, Dq);
            g.residual = x_in; g.ldr = D;
            D4_TRY(d4_engine_gemm(c, g, c->attn[i].w_out, 0, s));
        }
        D4_TRY(d4_row_rstd(hid(2 * i + 1), D, rowmap_identity(), M, D, hrs(2 * i + 1), s));
        D4_TRY(run_ff(c, c->ff[i], M, hid(2 * i + 1), D, rowmap_identity(), hrs(2 * i + 1), 0, hid(2 * i + 2), D, rowmap_identity(), nullptr, s));
        D4_TRY(d4_row_rstd(hid(2 * i + 2), D, rowmap_identity(), M, D, hrs(2 * i + 2), s));
        if (i != L - 1) {
            D4_TRY(run_pool(c, c->pools[i], M, hid(2 * i + 2), hrs(2 * i + 2), 0, 0, 2 * i + 3, c->b.x_cur, nullptr, s));
            D4_TRY(d4_row_rstd(c->b.x_cur, D, rowmap_identity(), M, D, xss(i), s));
            x_in = c->b.x_cur; x_in_rstd = xss(i);
        }
    }

    // ---- the special tokens cross-attend to the others once more, then their feed-forward (reference dreamer4.py:3227-3238)
    float* xf = c->b.x_cur;
    D4_CUDA_OK(cudaMemcpyAsync(xf, hid(2 * L), MD * 4, cudaMemcpyDeviceToDevice, s));
    const RowMap sp_rows = rowmap(ns, S, S - ns);
    D4_TRY(d4_row_rstd(xf, D, sp_rows, Ms, D, c->tfb.sp_rstd, s));
    {
        GemmArgs g = gemm_args(xf, D, nullptr, D, c->tfb.fa_q, c->ldfa, Ms, Dq + hq, D);
        g.amap = sp_rows; g.row_scale = c->tfb.sp_rstd;
        D4_TRY(d4_engine_gemm(c, g, c->fa.w_qg, 0, s));
        GemmArgs g2 = gemm_args(hid(2 * L), D, nullptr, D, c->b.fa_kv, 2 * Dkv, M, 2 * Dkv, D);
        g2.row_scale = hrs(2 * L);
        D4_TRY(d4_engine_gemm(c, g2, c->fa.w_kv, 0, s));
        SmallAttnArgs a; memset(&a, 0, sizeof(a));
        a.nb = B; a.hkv = h; a.g = hq / h; a.d = d; a.nq = ns; a.n = S - ns;
        a.q = c->tfb.fa_q; a.q_sb = (long long)ns * c->ldfa; a.q_si = c->ldfa;
        a.k = c->b.fa_kv; a.k_sb = (long long)S * 2 * Dkv; a.k_sj = 2 * Dkv;
        a.v = c->b.fa_kv + Dkv; a.v_sb = a.k_sb; a.v_sj = a.k_sj;
        a.k_gamma = c->fa.k_gamma;
        a.gate = c->tfb.fa_q + Dq; a.gate_sb = a.q_sb; a.gate_si = c->ldfa;
        a.out = c->tfb.fa_att; a.out_sb = (long long)ns * Dq; a.out_si = Dq;
        a.scale = att_scale;
        D4_TRY(attend(a));
        GemmArgs g3 = gemm_args(c->tfb.fa_att, Dq, nullptr, Dq, xf, D, Ms, D, Dq);
        g3.residual = xf; g3.ldr = D; g3.cmap = sp_rows;
        D4_TRY(d4_engine_gemm(c, g3, c->fa.w_out, 0, s));
    }
    D4_TRY(d4_row_rstd(xf, D, sp_rows, Ms, D, c->tfb.sp_rstd, s));
    D4_TRY(run_ff(c, c->fa_ff, Ms, xf, D, sp_rows, c->tfb.sp_rstd, 0, xf, D, sp_rows, nullptr, s));
    // final attention-residual pool over all 2L+1 hiddens (reference dreamer4.py:3242-3243), final norm (3247)
    D4_TRY(d4_row_rstd(xf, D, rowmap_identity(), M, D, xss(L), s));
    if (c->tf_final_norm) {
        D4_TRY(run_pool(c, c->pool_final, M, xf, xss(L), 0, 0, c->n_hid, xf, nullptr, s));
        return d4_rmsnorm_rows(xf, D, rowmap_identity(), c->tf_final_norm_w, M, D, tokens_out, D, s);
    }
    return run_pool(c, c->pool_final, M, xf, xss(L), 0, 0, c->n_hid, tokens_out, nullptr, s);
}

// ---- stand-alone operators of the tokenizer's front / back end
extern "C" int d4_linear_rows(int precision, int M, int N, int K, const float* A, int64_t lda, int a_grp, int a_gstride, int a_goff,
                              const float* W, int64_t ldw, const float* W_lo, const float* W_exact, const float* bias, float* C, int64_t ldc,
                              void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    GemmArgs g = gemm_args(A, lda, W_exact ? W_exact : W, ldw, C, ldc, M, N, K);
    g.bias = bias;
    if (a_grp > 0) g.amap = rowmap(a_grp, a_gstride, a_goff);
    if (precision == D4_PREC_FP32 || N < 16 || !d4_gemm_tc_supported(g)) {      // tiny N: a tensor-core tile would be > 90 % padding
        if (!W_exact) return d4_fail("d4_linear_rows: this shape runs on the exact-fp32 kernel and needs W_exact");
        return d4_gemm_simt(g, s);
    }
    if (precision == D4_PREC_TF32X3 || precision == D4_PREC_F16X3) {
        if (!W || !W_lo) return d4_fail("d4_linear_rows: tf32x3 needs the hi / lo words of W");
        g.W = W; g.W_lo = W_lo;
        return d4_gemm_tc(g, 3, s);
    }
    return d4_gemm_tc(g, 1, s);
}
extern "C" int d4_patchify(int B, int C, int H, int W, int p, const float* frame, int64_t stride_b, int64_t stride_c, float* out, void* stream) {
    return d4_patchify_launch(B, C, H, W, p, frame, stride_b, stride_c, out, static_cast<cudaStream_t>(stream));
}
extern "C" int d4_unpatchify_flow(int B, int C, int H, int W, int p, const float* patches, float* frame, int64_t stride_b, int64_t stride_c,
                                  float scale, void* stream) {
    return d4_unpatchify_flow_launch(B, C, H, W, p, patches, frame, stride_b, stride_c, scale, static_cast<cudaStream_t>(stream));
}
extern "C" int d4_tok_assemble(int B, int S, int P, int D, const float* lin, const float* ln_w, const float* pos_emb, const float* special,
                               int64_t special_bstride, int num_special, float* tokens, void* stream) {
    if (P + num_special != S) return d4_fail("d4_tok_assemble: %d patches + %d special tokens != %d tokens per frame", P, num_special, S);
    return d4_tok_assemble_launch(B, S, P, D, lin, ln_w, pos_emb, special, special_bstride, tokens, static_cast<cudaStream_t>(stream));
}
extern "C" int d4_tanh_rows(float* x, int64_t n, void* stream) { return d4_tanh_launch(x, n, static_cast<cudaStream_t>(stream)); }
