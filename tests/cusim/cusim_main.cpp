// TEST INFRASTRUCTURE - the simulator core (one OS thread per CUDA thread of a block, blocks one after the other) and host
// stand-ins for the kernels that are written in PTX and therefore outside the simulator (tensor-core GEMMs, the mma.sync /
// cp.async.bulk attention kernels): the engine under test is configured to exact-fp32 mode, where its GEMMs go to the SIMT
// kernel (simulated from its real source), and the attention entry points below restate the kernels' CONTRACT in plain loops.
// Linked with the transformed rowops / gemm_simt / frame_attn / tokenizer sources and engine.cu compiled as C++.
#include <float.h>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>
#include "engine.h"

thread_local dim3 threadIdx, blockIdx;
dim3 blockDim, gridDim;
namespace cusim {
thread_local Warp* warp = nullptr;
thread_local std::barrier<>* block_bar = nullptr;
alignas(16) unsigned char dyn_smem[1 << 20];

Graph* capturing = nullptr;

static void run(dim3 grid, dim3 block, const std::function<void()>& body);
void launch(dim3 grid, dim3 block, size_t smem, cudaStream_t, std::function<void()> body) {
    if (smem > sizeof(dyn_smem)) abort();
    enqueue([=] { run(grid, block, body); });
}
// A persistent pool of OS threads (one per CUDA thread of the largest block seen so far): per block the dispatcher publishes a job,
// the first `active` pool threads run the kernel body as CUDA threads 0..active-1, the dispatcher waits for all of them.
struct Pool {
    std::mutex m; std::condition_variable cv_start, cv_done;
    uint64_t gen = 0; unsigned active = 0, remaining = 0, bx = 0, by = 0;
    const std::function<void()>* body = nullptr; std::vector<Warp>* warps = nullptr; std::barrier<>* bb = nullptr;
    unsigned size = 0;
    void ensure(unsigned n) {
        std::unique_lock<std::mutex> lk(m);
        while (size < n) {
            const unsigned t = size++; const uint64_t seen0 = gen;
            std::thread([this, t, seen0] {
                uint64_t seen = seen0;
                for (;;) {
                    std::unique_lock<std::mutex> lk2(m);
                    cv_start.wait(lk2, [&] { return gen != seen; });
                    seen = gen;
                    if (t >= active) continue;
                    const std::function<void()>* fn = body; Warp* w = &(*warps)[t / 32]; std::barrier<>* blockbar = bb;
                    const unsigned x = bx, y = by;
                    lk2.unlock();
                    threadIdx = dim3(t); blockIdx = dim3(x, y);
                    warp = w; block_bar = blockbar;
                    (*fn)();
                    w->bar->arrive_and_drop();         // a thread that returned no longer takes part in barriers
                    blockbar->arrive_and_drop();
                    lk2.lock();
                    if (--remaining == 0) cv_done.notify_one();
                }
            }).detach();
        }
    }
    void run_block(unsigned nthreads, unsigned x, unsigned y, const std::function<void()>& fn, std::vector<Warp>& w, std::barrier<>& blockbar) {
        std::unique_lock<std::mutex> lk(m);
        active = remaining = nthreads; bx = x; by = y; body = &fn; warps = &w; bb = &blockbar;
        ++gen;
        cv_start.notify_all();
        cv_done.wait(lk, [&] { return remaining == 0; });
    }
};
static Pool* pool = new Pool();                        // never destroyed: its threads outlive static destruction

static void run(dim3 grid, dim3 block, const std::function<void()>& body) {
    gridDim = grid; blockDim = block;
    const unsigned nthreads = block.x, nwarps = (nthreads + 31) / 32;
    pool->ensure(nthreads);
    for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx) {
            std::barrier<> bb((std::ptrdiff_t)nthreads);
            std::vector<std::unique_ptr<std::barrier<>>> wbars;
            std::vector<Warp> warps(nwarps);
            for (unsigned w = 0; w < nwarps; ++w) {
                const unsigned lanes = (w + 1) * 32 <= nthreads ? 32 : nthreads - w * 32;
                wbars.emplace_back(new std::barrier<>((std::ptrdiff_t)lanes));
                warps[w].bar = wbars[w].get();
            }
            pool->run_block(nthreads, bx, by, body, warps, bb);
        }
}
}  // namespace cusim

// ---- tensor-core GEMMs: never reached in exact-fp32 mode
int d4_gemm_tc_supported(const GemmArgs&) { return 0; }
int d4_gemm_pair_default(void) { return 0; }
int d4_gemm_pair_bn(int, int) { return 256; }
int d4_gemm_tc(const GemmArgs&, int, cudaStream_t) { return d4_fail("cusim: tensor-core GEMM"); }
int d4_gemm_tc2(const GemmArgs&, int, int, cudaStream_t) { return d4_fail("cusim: tensor-core GEMM"); }
int d4_gemm_tc3(const GemmArgs&, int, int, cudaStream_t) { return d4_fail("cusim: tensor-core GEMM"); }
// ---- gemm_f16.cu's CONTRACT (the kernel itself is tcgen05 PTX): fp16 hi / lo words of q W, rows of A optionally pre-scaled by a
// power of two and split into fp16 hi / lo, three of the four cross products, fp32 result scaled by rs / p / q.  Lets the engine's
// f16x3 mode (weight registration, scales, dispatch, epilogue arguments) run end to end on the simulator.
static long long f16_calls = 0;
extern "C" long long sim_f16_calls(void) { return f16_calls; }
void d4_gemm_f16_debug(int) {}
int d4_gemm_f16x3_supported(const GemmArgs& g, const void* whi, const void* wlo) {
    return (whi && wlo && g.K % 8 == 0 && g.ldw % 8 == 0 && !g.transA && !g.transW && g.act != D4_ACT_SILU) ? 1 : 0;
}
static inline float pow2_near_h(float x) { return __uint_as_float((__float_as_uint(x) + 0x00400000u) & 0x7F800000u); }
int d4_gemm_f16x3(const GemmArgs& g0, float w_scale, int, cudaStream_t) {
    ++f16_calls;
    const GemmArgs g = g0;
    cusim::enqueue([g, w_scale] {
        const _Float16* whi = reinterpret_cast<const _Float16*>(g.W); const _Float16* wlo = reinterpret_cast<const _Float16*>(g.W_lo);
        const bool glu = g.act == D4_ACT_GLU_SILU || g.act == D4_ACT_GLU_GELU;
        std::vector<float> acc(g.N);
        for (int m = 0; m < g.M; ++m) {
            float p = 1.f, rs = g.row_scale ? g.row_scale[m] : 1.f;
            if (g.rs_mode && g.row_scale) { rs = 1.0f / sqrtf(rs / (float)g.K + D4_RMS_EPS); p = pow2_near_h(rs); rs /= p; }
            rs *= w_scale;
            const float* a = g.A + g.amap(m) * g.lda;
            for (int n = 0; n < g.N; ++n) {
                double sum = 0.0;
                for (int k = 0; k < g.K; ++k) {
                    const float x = a[k] * p;
                    const _Float16 ah = (_Float16)x; const _Float16 al = (_Float16)(x - (float)ah);
                    const float wh = (float)whi[(long long)n * g.ldw + k], wl = (float)wlo[(long long)n * g.ldw + k];
                    sum += (double)(float)al * wh + (double)(float)ah * wl + (double)(float)ah * wh;
                }
                acc[n] = (float)sum * rs + (g.bias ? g.bias[n] : 0.f);
            }
            const long long cr = g.cmap(m);
            const int nout = glu ? g.N / 2 : g.N;
            for (int j = 0; j < nout; ++j) {
                float v = acc[j];
                if (glu) { const float x = acc[2 * j], gt = acc[2 * j + 1]; v = x * (g.act == D4_ACT_GLU_SILU ? gt / (1.f + expf(-gt)) : 0.5f * gt * (1.f + erff(gt * 0.70710678f))); }
                if (g.residual) v += g.residual[cr * g.ldr + j];
                g.C[cr * g.ldc + j] = v;
            }
        }
    });
    return 0;
}
// ---- fused pools: report "unsupported", the engine then takes its GEMM + attention path
int d4_l2s_fused_supported(const L2sArgs&) { return 0; }
int d4_l2s_fused(const L2sArgs&, cudaStream_t) { return d4_fail("cusim: fused pool"); }
int d4_lp_fused_supported(const LpArgs&) { return 0; }
int d4_lp_fused(const LpArgs&, cudaStream_t) { return d4_fail("cusim: fused pool"); }
void d4_lp_fused_debug(int) {}

static inline float sigm(float x) { return 1.0f / (1.0f + expf(-x)); }
static inline float lerp_t(float a, float b, float w) { const float d = b - a; return (w < 0.5f) ? a + w * d : b - d * (1.f - w); }

// ---- attn.cu's contract (SmallAttnArgs in kernels.h): softmax(q . keynorm(k) * scale [softclamp] [agent mask]) @ lerp(v, v0, sigmoid(mix)),
// belief projection, head gate.  In-kernel gate logits (gate_w) are declined through d4_pool_attn_ok.
int d4_pool_attn_ok(const SmallAttnArgs&) { return 0; }
int d4_frame_attn_mma_ok(const SmallAttnArgs&) { return 0; }          // mma.sync tiles (frame_attn_mma.cu): hardware only
int d4_frame_attn_mma(const SmallAttnArgs&, cudaStream_t) { return d4_fail("cusim: frame_attn_mma"); }
static void small_attn_host(const SmallAttnArgs& a);
int d4_small_attn(const SmallAttnArgs& a, cudaStream_t) {
    if (a.gate_w) return d4_fail("cusim: in-kernel gate logits");
    const SmallAttnArgs copy = a;
    cusim::enqueue([copy] { small_attn_host(copy); });          // a launch: takes part in stream capture like one
    return 0;
}
static void small_attn_host(const SmallAttnArgs& a) {
    const int d = a.d, n = a.n;
    std::vector<float> K((size_t)n * d), V((size_t)n * d), p(n);
    for (int b = 0; b < a.nb; ++b)
        for (int hk = 0; hk < a.hkv; ++hk) {
            for (int j = 0; j < n; ++j) {
                const float* kr = a.k + b * a.k_sb + j * a.k_sj + (long long)hk * d;
                const float* vr = a.v + b * a.v_sb + j * a.v_sj + (long long)hk * d;
                float ss = 0.f;
                for (int c = 0; c < d; ++c) ss += kr[c] * kr[c];
                const float den = fmaxf(sqrtf(ss), 1e-12f);
                for (int c = 0; c < d; ++c) {
                    K[(size_t)j * d + c] = kr[c] / den * ((a.k_gamma[hk * d + c] + 1.f) * sqrtf((float)d));
                    float v = vr[c];
                    if (a.v0) v = lerp_t(v, a.v0[b * a.v0_sb + j * a.v0_sj + (long long)hk * d + c], sigm(a.mix[b * a.mix_sb + j * a.mix_sj + hk]));
                    V[(size_t)j * d + c] = v;
                }
            }
            for (int gi = 0; gi < a.g; ++gi) {
                const int hq = hk * a.g + gi;
                for (int i = 0; i < a.nq; ++i) {
                    const float* q = a.q + b * a.q_sb + i * a.q_si + (long long)hq * d;
                    float mx = -INFINITY;
                    for (int j = 0; j < n; ++j) {
                        float s = 0.f;
                        for (int c = 0; c < d; ++c) s += q[c] * K[(size_t)j * d + c];
                        s *= a.scale;
                        if (a.softclamp > 0.f) s = tanhf(s / a.softclamp) * a.softclamp;
                        if (a.mask_agent && i < a.nq - 1 && j == n - 1) s = -FLT_MAX;
                        p[j] = s; mx = fmaxf(mx, s);
                    }
                    float sum = 0.f;
                    for (int j = 0; j < n; ++j) { p[j] = expf(p[j] - mx); sum += p[j]; }
                    std::vector<float> o(d, 0.f);
                    for (int j = 0; j < n; ++j) for (int c = 0; c < d; ++c) o[c] += p[j] / sum * V[(size_t)j * d + c];
                    if (a.belief) {
                        float ss = 0.f, dot = 0.f;
                        for (int c = 0; c < d; ++c) ss += V[(size_t)i * d + c] * V[(size_t)i * d + c];
                        const float den = fmaxf(sqrtf(ss), 1e-12f);
                        for (int c = 0; c < d; ++c) dot += o[c] * V[(size_t)i * d + c] / den;
                        for (int c = 0; c < d; ++c) o[c] -= dot * V[(size_t)i * d + c] / den;
                    }
                    const float gate = a.gate ? sigm(a.gate[b * a.gate_sb + i * a.gate_si + hq]) : 1.f;
                    float* op = a.out + b * a.out_sb + i * a.out_si + (long long)hq * d;
                    for (int c = 0; c < d; ++c) op[c] = o[c] * gate;
                }
            }
        }
}

// ---- K1's contract (TimeAttnArgs in kernels.h): one new query per (token row, head) over the row's cached keys / values + itself;
// key = rope(keynorm(k)), query = rope(q) at position t, value = lerp(v, v0, sigmoid(mix)); appended at position t when commit.
static void time_attn_host(const TimeAttnArgs& a);
int d4_time_attn(const TimeAttnArgs& a, cudaStream_t) {
    const TimeAttnArgs copy = a;
    cusim::enqueue([copy] { time_attn_host(copy); });
    return 0;
}
static void time_attn_host(const TimeAttnArgs& a) {
    const int d = a.d, half = d / 2, t = a.t;
    std::vector<float> kn(d), vn(d), qr(d), p(t + 1), tmp(d);
    auto rope = [&](float* x) {
        for (int c = 0; c < d; ++c) tmp[c] = x[c];
        for (int c = 0; c < d; ++c) {
            const float ang = (float)t * a.inv_freq[c % half];
            const float rot = (c < half) ? -tmp[c + half] : tmp[c - half];
            x[c] = tmp[c] * cosf(ang) + rot * sinf(ang);
        }
    };
    for (int mc = 0; mc < a.M; ++mc) {
        const long long m = a.tmap(mc);                 // the launch may cover a subset of the token rows (TimeAttnArgs::tmap)
        const float* row = a.qkvgm + m * a.ld;
        for (int hk = 0; hk < a.hkv; ++hk) {
            const float* k = row + a.off_k + hk * d; const float* v = row + a.off_v + hk * d;
            float ss = 0.f;
            for (int c = 0; c < d; ++c) ss += k[c] * k[c];
            const float den = fmaxf(sqrtf(ss), 1e-12f);
            const float w = sigm(row[a.off_m + hk]);
            for (int c = 0; c < d; ++c) {
                kn[c] = k[c] / den * ((a.k_gamma[hk * d + c] + 1.f) * sqrtf((float)d));
                vn[c] = lerp_t(v[c], a.v0[(long long)m * a.ldv0 + hk * d + c], w);
            }
            rope(kn.data());
            float* kc = a.kcache + ((long long)m * a.hkv + hk) * a.Tmax * d;
            float* vc = a.vcache + ((long long)m * a.hkv + hk) * a.Tmax * d;
            for (int gi = 0; gi < a.g; ++gi) {
                const int hq = hk * a.g + gi;
                for (int c = 0; c < d; ++c) qr[c] = row[hq * d + c];
                rope(qr.data());
                float mx = -INFINITY;
                for (int j = 0; j <= t; ++j) {
                    const float* kj = (j < t) ? kc + (long long)j * d : kn.data();
                    float s = 0.f;
                    for (int c = 0; c < d; ++c) s += qr[c] * kj[c];
                    s *= a.scale;
                    if (a.softclamp > 0.f) s = tanhf(s / a.softclamp) * a.softclamp;
                    p[j] = s; mx = fmaxf(mx, s);
                }
                float sum = 0.f;
                for (int j = 0; j <= t; ++j) { p[j] = expf(p[j] - mx); sum += p[j]; }
                std::vector<float> o(d, 0.f);
                for (int j = 0; j <= t; ++j) { const float* vj = (j < t) ? vc + (long long)j * d : vn.data(); for (int c = 0; c < d; ++c) o[c] += p[j] / sum * vj[c]; }
                float vs = 0.f, dot = 0.f;
                for (int c = 0; c < d; ++c) vs += vn[c] * vn[c];
                const float vden = fmaxf(sqrtf(vs), 1e-12f);
                for (int c = 0; c < d; ++c) dot += o[c] * vn[c] / vden;
                const float gate = sigm(row[a.off_g + hq]);
                float* op = a.out + (long long)m * a.ldo + hq * d;
                for (int c = 0; c < d; ++c) op[c] = (o[c] - dot * vn[c] / vden) * gate;
            }
            if (a.commit) for (int c = 0; c < d; ++c) { kc[(long long)t * d + c] = kn[c]; vc[(long long)t * d + c] = vn[c]; }
        }
    }
}

// ---- a direct entry for the one simulated kernel the C-ABI does not expose on its own
// rowops.cu's LayerNorm + activation rows (head MLP hidden layers) for the unit test: register-resident and scalar variants by shape
extern "C" int sim_ln_act_rows(const float* x, long long ldx, const float* w, const float* b, int M, int D, float* out, long long ldo, int act,
                               float* save_mean, float* save_rstd) {
    return d4_ln_act_rows(x, ldx, w, b, M, D, out, ldo, act, save_mean, save_rstd, nullptr);
}

extern "C" int sim_frame_attn(int nb, int hkv, int g, int d, int nq, int n, const float* q, long long q_sb, long long q_si, const float* k,
                              long long k_sb, long long k_sj, const float* v, long long v_sb, long long v_sj, const float* k_gamma, const float* v0,
                              long long v0_sb, long long v0_sj, const float* mix, long long mix_sb, long long mix_sj, const float* gate,
                              long long gate_sb, long long gate_si, float* out, long long out_sb, long long out_si, float scale, float softclamp,
                              int num_special, int belief) {
    SmallAttnArgs a; memset(&a, 0, sizeof(a));
    a.nb = nb; a.hkv = hkv; a.g = g; a.d = d; a.nq = nq; a.n = n;
    a.q = q; a.q_sb = q_sb; a.q_si = q_si; a.k = k; a.k_sb = k_sb; a.k_sj = k_sj; a.v = v; a.v_sb = v_sb; a.v_sj = v_sj;
    a.k_gamma = k_gamma; a.v0 = v0; a.v0_sb = v0_sb; a.v0_sj = v0_sj; a.mix = mix; a.mix_sb = mix_sb; a.mix_sj = mix_sj;
    a.gate = gate; a.gate_sb = gate_sb; a.gate_si = gate_si; a.out = out; a.out_sb = out_sb; a.out_si = out_si;
    a.scale = scale; a.softclamp = softclamp; a.mask_agent = num_special; a.belief = belief;
    return d4_frame_attn(a, nullptr);
}
