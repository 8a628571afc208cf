"""The parts of the native library that are plain CUDA C - the engine's orchestration (engine.cu), the row kernels, the SIMT
GEMM, and the kernels drafted in round 1 without a GPU at hand (frame_attn.cu, tokenizer.cu) - compiled for the HOST by g++ and
executed under a thread-per-CUDA-thread simulator (tests/cusim/: grids, blocks, shared memory, __syncthreads / __syncwarp, warp
shuffles; `<<<...>>>` launches rewritten by tests/cusim/transform.py).  The kernels written in PTX (tensor-core GEMMs, K1, the
mma.sync attention) are outside the simulator: the engine runs in exact-fp32 mode, where its GEMMs use the SIMT kernel, and the
attention entry points are restated as plain loops from their contract (tests/cusim/cusim_main.cpp).

What this establishes on the CPU: the engine's buffer plumbing, strides and row maps, and the indexing / masking / reductions /
barriers of the simulated kernels - for the hardware-verified dynamics pass (which validates the harness itself against the
oracle) and for the tokenizer path that has not run on hardware yet (d4_tf_step and the VideoTokenizer host class end to end,
against the REFERENCE's golden vectors).  Not performance, not memory-model subtleties, not the PTX kernels: the -m gpu tests
remain the parity tests proper.  Test infrastructure only - nothing here is linked into the product."""
import contextlib
import ctypes as C
import glob
import os
import subprocess
import sys

import pytest
import torch

from engine_emulator import patchify, small_attn, tok_assemble, unpatchify
from oracle import dreamer4_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), 'dreamer4_b200', 'csrc')
SIM = os.path.join(HERE, 'cusim')
sys.path.insert(0, SIM)
LL = C.c_longlong


@pytest.fixture(scope='module')
def simlib(tmp_path_factory):
    from transform import transform
    from dreamer4_b200 import _lib
    build = tmp_path_factory.mktemp('cusim')
    srcs = []
    for name in ('engine', 'learn', 'rowops', 'gemm_simt', 'gemm_skinny', 'frame_attn', 'tokenizer'):
        out = build / f'{name}.cpp'
        out.write_text(transform(open(os.path.join(CSRC, name + '.cu')).read()))
        srcs.append(str(out))
    lib = build / 'libcusim.so'
    extra = ['-fsanitize=address', '-fno-omit-frame-pointer', '-g'] if os.environ.get('D4_CUSIM_ASAN') == '1' else []      # see tests/cusim/README.md
    cmd = ['g++', '-std=c++20', '-O2', *extra, '-shared', '-fPIC', '-pthread', '-I', SIM, '-I', CSRC, *srcs, os.path.join(SIM, 'cusim_main.cpp'), '-o', str(lib)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    lib = C.CDLL(str(lib))
    for name, (res, args) in _lib.SYMBOLS.items():          # the product's own C-ABI, where the simulated sources define it
        if hasattr(lib, name):
            getattr(lib, name).restype, getattr(lib, name).argtypes = res, args
    return lib


class _Stream:
    cuda_stream = 0


@pytest.fixture
def on_simulator(simlib, monkeypatch):
    """Routes the host classes' native calls to the simulated library and lifts their CUDA-only guards."""
    from dreamer4_b200 import DynamicsWorldModel, VideoTokenizer, _lib
    monkeypatch.setattr(_lib, '_lib', simlib)
    # the <= 32-row weight-streaming GEMM (one 128-thread block per 4 weight rows, 5-step shuffle trees) is slow to SIMULATE thread by
    # thread: the engine-level tests keep the tile kernel, test_skinny_gemm_on_the_simulator below covers the kernel itself
    monkeypatch.setenv('D4_SKINNY', '0')
    for cls in (DynamicsWorldModel, VideoTokenizer):
        monkeypatch.setattr(cls, '_require_cuda', lambda self: None)
    monkeypatch.setattr(VideoTokenizer, '_stream', lambda self: C.c_void_p(0))
    monkeypatch.setattr(torch.cuda, 'current_stream', lambda device=None: _Stream())
    monkeypatch.setattr(torch.cuda, 'device', lambda device=None: contextlib.nullcontext())
    monkeypatch.setattr(torch.Tensor, 'record_stream', lambda self, stream: None)
    return simlib


def p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


# ------------------------------------------------------------------------------------------------ gemm_skinny.cu

@pytest.mark.parametrize('M,N,K,act,res', [(1, 8, 16, 0, False), (15, 40, 64, 0, True), (7, 24, 32, 1, False), (32, 10, 48, 0, True), (20, 6, 36, 2, False)])
def test_skinny_gemm_on_the_simulator(simlib, M, N, K, act, res):
    """d4_linear(fp32) with <= 32 rows takes gemm_skinny.cu: every row-count template, ragged N, GLU pairs, bias / row scale / residual."""
    torch.manual_seed(M * 100 + N)
    A, W = torch.randn(M, K), torch.randn(N, K) / K ** 0.5
    bias, rs = torch.randn(N), torch.rand(M) + 0.5
    nout = N // 2 if act else N
    R = torch.randn(M, nout) if res else None
    Cc = torch.full((M, nout), float('nan'))
    rc = simlib.d4_linear(0, M, N, K, p(A), K, p(W), K, None, p(bias), p(rs), p(R), nout, act, p(Cc), nout, None)
    assert rc == 0, simlib.d4_last_error()
    ref = (A.double() @ W.double().T) * rs.double()[:, None] + bias.double()
    if act:
        x, g = ref[:, 0::2], ref[:, 1::2]
        ref = x * (torch.nn.functional.silu(g) if act == 1 else torch.nn.functional.gelu(g))
    elif res:
        ref = ref + R.double()
    torch.testing.assert_close(Cc.double(), ref, atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize('M,N,K', [(256, 4, 128), (300, 8, 64), (257, 1, 36)])
def test_rowdot_gemm_on_the_simulator(simlib, M, N, K):
    """d4_linear(fp32) with many rows and at most 8 output columns, no epilogue (the policy unembedding) takes gemm_simt.cu's
    one-warp-per-row kernel: ragged row count, K not a multiple of the 128-float lane stride."""
    torch.manual_seed(M + N)
    A, W = torch.randn(M, K), torch.randn(N, K) / K ** 0.5
    Cc = torch.full((M, N), float('nan'))
    rc = simlib.d4_linear(0, M, N, K, p(A), K, p(W), K, None, None, None, None, N, 0, p(Cc), N, None)
    assert rc == 0, simlib.d4_last_error()
    torch.testing.assert_close(Cc.double(), A.double() @ W.double().T, atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize('M,D', [(9, 512), (5, 2048), (3, 100), (4, 2052)])
def test_layernorm_silu_rows_on_the_simulator(simlib, M, D):
    """d4_mlp-style hidden layer rows: LayerNorm + SiLU with the row held in registers (D <= 2048, D % 4 == 0) and the scalar fallback."""
    if not hasattr(simlib, 'sim_ln_act_rows'):
        pytest.skip('simulator build without the ln_act_rows entry point')
    torch.manual_seed(D)
    x, w, b = torch.randn(M, D) * 3 + 1, torch.randn(D), torch.randn(D)
    out, mean, rstd = torch.empty(M, D), torch.empty(M), torch.empty(M)
    f = simlib.sim_ln_act_rows
    f.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p]
    f.restype = C.c_int
    assert f(p(x), D, p(w), p(b), M, D, p(out), D, 3, p(mean), p(rstd)) == 0
    ref = torch.nn.functional.silu(torch.nn.functional.layer_norm(x.double(), (D,), w.double(), b.double(), 1e-5))
    torch.testing.assert_close(out.double(), ref, atol=2e-5, rtol=2e-5)
    torch.testing.assert_close(mean.double(), x.double().mean(1), atol=1e-5, rtol=1e-5)


# ------------------------------------------------------------------------------------------------ frame_attn.cu

def frame_attn(sim, q, k, v, k_gamma, scale, v0=None, mix=None, gate=None, softclamp=0., num_special=0, belief=False):
    """q (b, nq, hq, d); k, v (b, n, h, d) contiguous; returns (b, nq, hq, d)."""
    b, nq, hq, d = q.shape
    n, h = k.shape[1], k.shape[2]
    out = torch.full((b, nq, hq, d), float('nan'))
    f = sim.sim_frame_attn
    f.argtypes = [C.c_int] * 6 + [C.c_void_p, LL, LL] * 3 + [C.c_void_p] + [C.c_void_p, LL, LL] * 4 + [C.c_float, C.c_float, C.c_int, C.c_int]
    rc = f(b, h, hq // h, d, nq, n, p(q), nq * hq * d, hq * d, p(k), n * h * d, h * d, p(v), n * h * d, h * d, p(k_gamma),
           p(v0), n * h * d, h * d, p(mix), n * h, h, p(gate), nq * hq, hq, p(out), nq * hq * d, hq * d, scale, softclamp, num_special, int(belief))
    assert rc == 0
    return out


@pytest.mark.parametrize('S,ns,h,g,d', [(22, 6, 2, 1, 16), (70, 1, 2, 2, 32), (40, 8, 1, 1, 64), (33, 32, 2, 1, 16)])
def test_frame_attn_self_attention(simlib, S, ns, h, g, d):
    """Space attention of one frame: key RMSNorm, softclamp, special-token mask, value-residual lerp, belief projection, gates."""
    torch.manual_seed(S + ns)
    b, hq = 2, h * g
    q, k, v, v0 = torch.randn(b, S, hq, d), torch.randn(b, S, h, d), torch.randn(b, S, h, d), torch.randn(b, S, h, d)
    mix, gate, gamma = torch.randn(b, S, h), torch.randn(b, S, hq), torch.randn(h, d) * 0.2
    got = frame_attn(simlib, q, k, v, gamma, d ** -0.5, v0=v0, mix=mix, gate=gate, softclamp=50., num_special=ns, belief=True)
    want, _, _ = small_attn(q, k, v, gamma, d ** -0.5, gate=gate, softclamp=50., num_special=ns, belief=True, v0=v0, mix=mix)
    torch.testing.assert_close(got, want, atol=2e-5, rtol=1e-4)


@pytest.mark.parametrize('nq,n,h,d', [(6, 16, 2, 16), (1, 69, 2, 32), (64, 64, 1, 64)])
def test_frame_attn_cross_attention(simlib, nq, n, h, d):
    """The special tokens' final cross-attention over the other tokens: no mask, no softclamp, no belief, gated."""
    torch.manual_seed(nq + n)
    b = 2
    q, k, v = torch.randn(b, nq, h, d), torch.randn(b, n, h, d), torch.randn(b, n, h, d)
    gate, gamma = torch.randn(b, nq, h), torch.randn(h, d) * 0.2
    got = frame_attn(simlib, q, k, v, gamma, d ** -0.5, gate=gate)
    want, _, _ = small_attn(q, k, v, gamma, d ** -0.5, gate=gate)
    torch.testing.assert_close(got, want, atol=2e-5, rtol=1e-4)


# ------------------------------------------------------------------------------------------------ tokenizer.cu through the C-ABI

def test_tokenizer_frame_ops(simlib):
    """d4_patchify / d4_tok_assemble / d4_unpatchify_flow / d4_tanh_rows / d4_linear_rows (the same calls as
    tests/test_zy_tokenizer_gpu.py::test_frame_ops_match_torch makes on the GPU)."""
    lib, s = simlib, C.c_void_p(0)
    torch.manual_seed(0)
    B, Cc, T, H, W, pp, D, N = 2, 3, 3, 8, 12, 4, 24, 5
    P = (H // pp) * (W // pp)
    video = torch.randn(B, Cc, T, H, W)
    frame = video[:, :, 1]
    out = torch.full((B * P, pp * pp * Cc), float('nan'))
    assert lib.d4_patchify(B, Cc, H, W, pp, p(frame), frame.stride(0), frame.stride(1), p(out), s) == 0
    assert torch.equal(out, patchify(frame, pp))
    lin, ln_w, pos, spec = torch.randn(B * P, D), torch.randn(D), torch.randn(P, D), torch.randn(B, N, D)
    tok = torch.full((B, P + N, D), float('nan'))
    assert lib.d4_tok_assemble(B, P + N, P, D, p(lin), p(ln_w), p(pos), p(spec), N * D, N, p(tok), s) == 0
    torch.testing.assert_close(tok, tok_assemble(lin, ln_w, pos, spec, B, P), atol=1e-5, rtol=1e-5)
    tok.fill_(float('nan'))
    assert lib.d4_tok_assemble(B, P + N, P, D, p(lin), p(ln_w), None, p(spec[0].contiguous()), 0, N, p(tok), s) == 0
    torch.testing.assert_close(tok, tok_assemble(lin, ln_w, None, spec[0], B, P), atol=1e-5, rtol=1e-5)
    assert lib.d4_tok_assemble(B, P + N, P, D, p(lin), p(ln_w), None, p(spec), 0, N + 1, p(tok), s) != 0           # P + num_special != S
    pred = torch.randn(B * P, pp * pp * Cc)
    want = frame + (unpatchify(pred, B, pp, Cc, H, W) - frame) * 0.75
    untouched = video[:, :, 0].clone()
    assert lib.d4_unpatchify_flow(B, Cc, H, W, pp, p(pred), p(frame), frame.stride(0), frame.stride(1), 0.75, s) == 0
    torch.testing.assert_close(video[:, :, 1], want, atol=1e-6, rtol=1e-6)
    assert torch.equal(video[:, :, 0], untouched)
    x = torch.randn(1000)
    want = x.tanh()
    assert lib.d4_tanh_rows(p(x), x.numel(), s) == 0
    torch.testing.assert_close(x, want, atol=1e-6, rtol=1e-6)
    S, K, Nn = P + N, 20, 7                       # rows of A through a grouped map: the N special rows of each of B frames of S tokens
    A, Wt, bias = torch.randn(B * S, K), torch.randn(Nn, K), torch.randn(Nn)
    Cout = torch.full((B * N, Nn), float('nan'))
    assert lib.d4_linear_rows(0, B * N, Nn, K, p(A), K, N, S, P, p(Wt), K, None, p(Wt), p(bias), p(Cout), Nn, s) == 0
    torch.testing.assert_close(Cout, A.view(B, S, K)[:, P:].reshape(B * N, K) @ Wt.T + bias, atol=1e-5, rtol=1e-5)


# ------------------------------------------------------------------------------------------------ the engine, end to end

GOLDEN = sorted(glob.glob(os.path.join(HERE, 'golden', '*.pt')))


@pytest.mark.parametrize('path', GOLDEN[2:], ids=[os.path.basename(x)[:-3] for x in GOLDEN[2:]])       # the other two architectures: graph / f16x3 tests below
def test_dynamics_rollout_on_the_simulator_matches_oracle(path, on_simulator):
    """The hardware-verified path first - DynamicsWorldModel.generate through the real engine.cu on simulated kernels against the
    oracle: this is what validates the harness (the stand-ins for the PTX kernels included)."""
    from dreamer4_b200 import DynamicsWorldModel
    fx = torch.load(path, map_location='cpu', weights_only=False)
    model = DynamicsWorldModel(**fx['model_kwargs'], precision='fp32')
    model.load_state_dict(fx['state_dict'], strict=True)
    ocfg = O.config_from_reference_kwargs(**fx['model_kwargs'])
    T, B = 2, 2
    g = torch.Generator().manual_seed(3)
    A = sum(model.cfg.num_discrete_actions)
    noise = dict(latent=torch.randn(T, B, model.cfg.num_latent_tokens, model.cfg.dim_latent, generator=g), action_uniform=torch.rand(T, B, A, generator=g),
                 terminal_uniform=torch.rand(T, B, generator=g))
    ref = O.generate(fx['state_dict'], ocfg, T, B, noise=O.InjectedNoise(noise['latent'], noise['action_uniform'], noise['terminal_uniform']))
    try:
        exp = model.generate(T, batch_size=B, noise=noise, return_rewards_per_frame=True, return_agent_actions=True, return_log_probs_and_values=True)
    finally:
        model._release()
    assert torch.equal(exp.actions.discrete, ref.actions)
    for name in ('latents', 'rewards', 'values', 'agent_embed'):
        torch.testing.assert_close(getattr(exp, name), getattr(ref, name), atol=5e-5, rtol=2e-4, msg=lambda m, n=name: f'{n}: {m}')


TOK = sorted(glob.glob(os.path.join(HERE, 'golden', 'tokenizer', 'tokenizer_*.pt')))


@pytest.mark.parametrize('path', TOK, ids=[os.path.basename(x)[:-3] for x in TOK])
def test_video_tokenizer_on_the_simulator_reproduces_reference_golden(path, on_simulator):
    """The path that has NOT run on hardware yet: VideoTokenizer.tokenize / .decode - host class, d4_tf_create / d4_tf_step, frame_attn.cu,
    tokenizer.cu, d4_linear_rows - against the vectors the reference's own source produced."""
    from dreamer4_b200 import VideoTokenizer
    fx = torch.load(path, map_location='cpu', weights_only=False)
    tok = VideoTokenizer(**fx['tokenizer_kwargs'], precision='fp32')
    tok.load_state_dict(fx['state_dict'], strict=True)
    try:
        latents = tok.tokenize(fx['video'])
        torch.testing.assert_close(latents, fx['latents'], atol=5e-5, rtol=2e-4)
        first, cache = tok.tokenize(fx['video'][:, :, :1], return_time_cache=True)                 # resumed over the encoder's time cache
        rest = tok.tokenize(fx['video'][:, :, 1:], time_cache=cache)
        torch.testing.assert_close(torch.cat((first, rest), dim=1), fx['latents'], atol=5e-5, rtol=2e-4)
        b, c, T, H, W = fx['video'].shape
        torch.manual_seed(fx['decode_seed'])
        noise = torch.randn(b, c, T, H, W)                                 # the draw at reference dreamer4.py:4204
        recon = tok.decode(fx['latents'], noise=noise)
        torch.testing.assert_close(recon, fx['recon'], atol=1e-4, rtol=2e-4)
    finally:
        tok._release()
    # the class default precision: every GEMM weight's tf32 hi / lo words must be registered under the names d4_bind resolves (on the
    # simulator the GEMMs themselves still take the exact kernel; the point is that binding and the d4_linear_rows arguments hold)
    tok3 = VideoTokenizer(**fx['tokenizer_kwargs'])
    assert tok3.precision == 'tf32x3'
    tok3.load_state_dict(fx['state_dict'], strict=True)
    try:
        torch.testing.assert_close(tok3.tokenize(fx['video'][:, :, :1]), fx['latents'][:, :1], atol=5e-5, rtol=2e-4)
    finally:
        tok3._release()


def test_cuda_graph_replay_on_the_simulator(on_simulator, monkeypatch):
    """engine.cu: frame_impl with D4_GRAPH=1 - first rollout direct, second captured (stream capture modelled as recording the
    enqueued launches / copies) and replayed over the staging rows, third replayed: identical Experiences, identical launch counts
    (what tests/test_zx_graph_replay_gpu.py asks of the hardware)."""
    from dreamer4_b200 import DynamicsWorldModel
    monkeypatch.setenv('D4_GRAPH', '1')
    fx = torch.load(GOLDEN[1], map_location='cpu', weights_only=False)         # GQA + two action types
    model = DynamicsWorldModel(**fx['model_kwargs'], precision='fp32')
    model.load_state_dict(fx['state_dict'], strict=True)
    T, B = 1, 2          # one cache position: seen directly, then captured, then replayed
    g = torch.Generator().manual_seed(3)
    A = sum(model.cfg.num_discrete_actions)
    noise = dict(latent=torch.randn(T, B, model.cfg.num_latent_tokens, model.cfg.dim_latent, generator=g), action_uniform=torch.rand(T, B, A, generator=g),
                 terminal_uniform=torch.rand(T, B, generator=g))
    runs, launches = [], []
    try:
        for _ in range(3):
            l0 = on_simulator.d4_launch_count()
            runs.append(model.generate(T, batch_size=B, noise=noise, return_rewards_per_frame=True, return_agent_actions=True, return_log_probs_and_values=True))
            launches.append(on_simulator.d4_launch_count() - l0)
    finally:
        model._release()
    assert launches[0] == launches[1] == launches[2] > 0
    for later in runs[1:]:
        assert torch.equal(later.actions.discrete, runs[0].actions.discrete)
        for name in ('latents', 'rewards', 'values', 'agent_embed', 'lens', 'terminals'):
            assert torch.equal(getattr(later, name), getattr(runs[0], name)), name
        assert torch.equal(later.log_probs.discrete, runs[0].log_probs.discrete)
        assert torch.equal(later.old_action_unembeds.discrete, runs[0].old_action_unembeds.discrete)


def test_f16x3_engine_mode_plumbing_on_the_simulator(on_simulator):
    """DynamicsWorldModel(precision='f16x3') through the real engine: fp16 hi / lo weights and their scales registered by the host,
    resolved by d4_bind, dispatched by d4_engine_gemm with the right epilogue arguments.  The fp16 kernel itself is tcgen05 PTX and
    is restated from its contract here (tests/cusim/cusim_main.cpp) - this test is about everything around it."""
    from dreamer4_b200 import DynamicsWorldModel
    on_simulator.sim_f16_calls.restype = C.c_longlong
    fx = torch.load(GOLDEN[0], map_location='cpu', weights_only=False)
    model = DynamicsWorldModel(**fx['model_kwargs'], precision='f16x3')
    model.load_state_dict(fx['state_dict'], strict=True)
    ocfg = O.config_from_reference_kwargs(**fx['model_kwargs'])
    T, B = 1, 2
    g = torch.Generator().manual_seed(3)
    A = sum(model.cfg.num_discrete_actions)
    noise = dict(latent=torch.randn(T, B, model.cfg.num_latent_tokens, model.cfg.dim_latent, generator=g), action_uniform=torch.rand(T, B, A, generator=g),
                 terminal_uniform=torch.rand(T, B, generator=g))
    ref = O.generate(fx['state_dict'], ocfg, T, B, noise=O.InjectedNoise(noise['latent'], noise['action_uniform'], noise['terminal_uniform']))
    before = on_simulator.sim_f16_calls()
    try:
        exp = model.generate(T, batch_size=B, noise=noise, return_rewards_per_frame=True, return_agent_actions=True, return_log_probs_and_values=True)
    finally:
        model._release()
    assert on_simulator.sim_f16_calls() - before > 50          # the transformer's dense layers really went through the fp16 entry point
    assert torch.equal(exp.actions.discrete, ref.actions)
    for name in ('latents', 'rewards', 'values', 'agent_embed'):
        torch.testing.assert_close(getattr(exp, name), getattr(ref, name), atol=5e-5, rtol=2e-4, msg=lambda m, n=name: f'{n}: {m}')


def test_learn_from_experience_on_the_simulator_matches_reference_golden(on_simulator):
    """d4_learn (learn.cu: GAE scan, the three heads' forward / backward, the policy row kernel) through the real engine on simulated
    kernels against the losses and gradients the reference's own source produced (the test tests/test_gpu_parity.py runs on the GPU)."""
    from dreamer4_b200 import Actions, DynamicsWorldModel, Experience
    fx = torch.load(GOLDEN[0], map_location='cpu', weights_only=False)
    model = DynamicsWorldModel(**fx['model_kwargs'], precision='fp32')
    model.load_state_dict(fx['state_dict'], strict=True)
    ref = fx['out']
    exp = Experience(latents=ref['latents'], agent_embed=ref['agent_embed'], rewards=ref['rewards'], values=ref['values'],
                     actions=Actions(ref['actions'], None), log_probs=Actions(ref['log_probs'], None), lens=ref['lens'],
                     is_truncated=ref['is_truncated'], terminals=ref['terminals'], step_size=ref['step_size'])
    try:
        pl, vl = model.learn_from_experience(exp)
        torch.testing.assert_close(pl.detach(), ref['policy_loss'], atol=1e-6, rtol=1e-4)
        torch.testing.assert_close(vl.detach(), ref['value_loss'], atol=1e-6, rtol=1e-4)
        pl.backward()
        vl.backward()
    finally:
        model._release()
    params = dict(model.named_parameters())
    for name, g in ref['grads'].items():
        assert params[name].grad is not None, name
        torch.testing.assert_close(params[name].grad, g, atol=2e-6, rtol=2e-4, msg=lambda m, n=name: f'{n}: {m}')


def test_reward_ema_stats_on_the_simulator_match_reference_golden(on_simulator):
    """keep_reward_ema_stats=True (reference dreamer4.py:5987-6013) through the real host class and learn.cu: two consecutive updates of
    the reference on one dream (oracle/make_golden_learn_ema.py); lens / is_truncated vary per dream."""
    from dreamer4_b200 import Actions, DynamicsWorldModel, Experience
    fx = torch.load(os.path.join(HERE, 'golden', 'learn', 'learn_ema.pt'), map_location='cpu', weights_only=False)
    model = DynamicsWorldModel(**fx['model_kwargs'], precision='fp32')
    model.load_state_dict(fx['state_dict'], strict=True)
    e = fx['experience']
    exp = Experience(latents=e['latents'], agent_embed=e['agent_embed'], rewards=e['rewards'], values=e['values'], actions=Actions(e['actions'], None),
                     log_probs=Actions(e['log_probs'], None), lens=e['lens'], is_truncated=e['is_truncated'], terminals=e['terminals'],
                     step_size=e['step_size'], old_action_unembeds=Actions(e['old_action_unembeds'], None))
    params = dict(model.named_parameters())
    try:
        for call in fx['calls']:
            model.zero_grad()
            pl, vl = model.learn_from_experience(exp, objective=call['objective'])
            torch.testing.assert_close(model.ema_returns_mean, call['ema_returns_mean'], atol=1e-6, rtol=1e-5)
            torch.testing.assert_close(model.ema_returns_var, call['ema_returns_var'], atol=1e-6, rtol=1e-5)
            torch.testing.assert_close(pl.detach(), call['policy_loss'], atol=1e-6, rtol=1e-4)
            torch.testing.assert_close(vl.detach(), call['value_loss'], atol=1e-6, rtol=1e-4)
            pl.backward()
            vl.backward()
            for name, g in call['grads'].items():
                torch.testing.assert_close(params[name].grad, g, atol=2e-6, rtol=2e-4, msg=lambda m, n=name: f'{n}: {m}')
    finally:
        model._release()


def test_forward_inference_branch_on_the_simulator_matches_reference_golden(on_simulator):
    """DynamicsWorldModel.forward (inference branch, reference dreamer4.py:6792-7295) through d4_pass_ex, and the head modules through
    d4_head_forward, against the reference's own numbers (oracle/make_golden_forward.py): parallel 4-frame call with per-dream signal
    levels / step sizes / actions / tasks, then the same frames one by one over the returned time cache."""
    from dreamer4_b200 import DynamicsWorldModel
    fx = torch.load(os.path.join(HERE, 'golden', 'forward', 'forward_inference.pt'), map_location='cpu', weights_only=False)
    model = DynamicsWorldModel(**fx['model_kwargs'], precision='fp32')
    model.load_state_dict(fx['state_dict'], strict=True)
    tol = dict(atol=5e-5, rtol=2e-4)
    kw = dict(step_sizes=fx['step_sizes'], tasks=fx['tasks'], return_pred_only=True, return_intermediates=True, latent_is_noised=True)
    try:
        pred, (embeds, inter) = model(latents=fx['latents'], signal_levels=fx['signal_levels'], discrete_actions=fx['actions'], **kw)
        torch.testing.assert_close(pred.flow, fx['flow'], **tol)
        torch.testing.assert_close(embeds.agent, fx['agent'], **tol)
        assert inter.main.token_count == fx['token_count']
        torch.testing.assert_close(inter.main.next_kv_cache, fx['kv_cache'], **tol)
        torch.testing.assert_close(model.policy_head(embeds.agent), fx['policy_embed'], **tol)
        torch.testing.assert_close(model.value_head(embeds.agent), fx['value_bins'], **tol)
        flows, agents, cache = [], [], None
        for i in range(fx['latents'].shape[1]):
            act = None if i == 0 else fx['actions'][:, i - 1:i]
            p_i, (e_i, cache) = model(latents=fx['latents'][:, i:i + 1], signal_levels=fx['signal_levels'][:, i:i + 1], discrete_actions=act, time_cache=cache, **kw)
            flows.append(p_i.flow.clone())
            agents.append(e_i.agent.clone())
        torch.testing.assert_close(torch.cat(flows, dim=1), fx['seq_flow'], **tol)
        torch.testing.assert_close(torch.cat(agents, dim=1), fx['seq_agent'], **tol)
        torch.testing.assert_close(cache.main.next_kv_cache, fx['seq_kv_cache'], **tol)
    finally:
        model._release()


# the two shortcut cases run three transformer passes thread-by-thread (4-6 min each on this box): on the simulator only with
# D4_CUSIM_SLOW=1; tests/test_forward_gpu.py holds all four cases to the same golden on hardware
_SLOW = pytest.mark.skipif(os.environ.get('D4_CUSIM_SLOW') != '1', reason='slow on the CPU simulator (D4_CUSIM_SLOW=1); covered on the GPU')


@pytest.mark.parametrize('case', [pytest.param(0, marks=_SLOW), pytest.param(1, marks=_SLOW), 2, 3], ids=['shortcut', 'shortcut_var_len', 'plain', 'plain_var_len'])
def test_training_forward_losses_on_the_simulator_match_reference_golden(on_simulator, case):
    """The TRAINING branch of DynamicsWorldModel.forward, forward only (reference dreamer4.py:6963-6997, 7297-7743): flow, shortcut,
    multi-token reward / action and terminal losses and their total against the reference's own numbers, fed the schedule and the noise
    the reference's seed produced (oracle/make_golden_training_forward.py)."""
    from dreamer4_b200 import DynamicsWorldModel
    fx = torch.load(os.path.join(HERE, 'golden', 'forward', 'forward_training.pt'), map_location='cpu', weights_only=False)
    ref = fx['cases'][case]
    model = DynamicsWorldModel(**fx['model_kwargs'], precision='fp32')
    model.load_state_dict(fx['state_dict'], strict=True)
    try:
        kw = dict(latents=fx['latents'], rewards=fx['rewards'], discrete_actions=fx['actions'], terminals=fx['terminals'], tasks=fx['tasks'], return_all_losses=True,
                  noise=ref['noise'], train_schedule=dict(shortcut_train=ref['shortcut_train'], step_sizes_log2=ref['step_sizes_log2'], signal_levels=ref['signal_levels']))
        if ref['var_len']:
            kw['lens'] = fx['lens']
        total, losses = model(**kw)
    finally:
        model._release()
    for name in ('flow', 'shortcut', 'rewards', 'terminals', 'discrete_actions'):
        torch.testing.assert_close(getattr(losses, name), ref['losses'][name], atol=2e-6, rtol=2e-4, msg=lambda m, n=name: f'{n}: {m}')
    torch.testing.assert_close(total, ref['total'], atol=1e-5, rtol=1e-4)
    assert bool(ref['shortcut_train']) == bool(float(losses.shortcut) > 0)


def test_trimmed_final_pool_equals_full_on_the_simulator(on_simulator, monkeypatch):
    """The final attention-residual pool (and the agent cross-attention) computed only for the token rows a pass's outputs read
    (engine.cu: run_pool_rows, D4_TRIM_FINAL) gives bit-identical rollouts to computing every row - the pool is per token."""
    from dreamer4_b200 import DynamicsWorldModel
    fx = torch.load(GOLDEN[1], map_location='cpu', weights_only=False)
    runs = []
    for trim in ('1', '0'):
        monkeypatch.setenv('D4_TRIM_FINAL', trim)
        monkeypatch.setenv('D4_TRIM_CONE', trim)
        model = DynamicsWorldModel(**fx['model_kwargs'], precision='fp32')
        model.load_state_dict(fx['state_dict'], strict=True)
        T, B = 3, 2
        g = torch.Generator().manual_seed(3)
        A = sum(model.cfg.num_discrete_actions)
        noise = dict(latent=torch.randn(T, B, model.cfg.num_latent_tokens, model.cfg.dim_latent, generator=g), action_uniform=torch.rand(T, B, A, generator=g),
                     terminal_uniform=torch.rand(T, B, generator=g))
        try:
            runs.append(model.generate(T, batch_size=B, return_rewards_per_frame=True, return_agent_actions=True, return_log_probs_and_values=True, noise=noise))
        finally:
            model._release()
    a, b = runs
    assert torch.equal(a.actions.discrete, b.actions.discrete)
    for name in ('latents', 'rewards', 'values', 'agent_embed'):
        assert torch.equal(getattr(a, name), getattr(b, name)), name


def test_sim_trainer_on_the_simulator(on_simulator):
    """SimTrainer (reference trainers.py:1472-1790) end to end on the toy image env: episodes through interact_with_env (d4_observe),
    combined, replayed in shuffled minibatches through d4_learn, both heads stepped.  The first minibatch's losses equal a direct
    learn_from_experience call on the same rows; only the policy / value heads (and the action unembedding) move."""
    from torch.utils.data import DataLoader, TensorDataset
    from dreamer4_b200 import Actions, DynamicsWorldModel, Experience, SimTrainer, combine_experiences
    from oracle import tokenizer_oracle as TO
    from oracle.toy_env import ToyImageEnv
    fx = torch.load(os.path.join(HERE, 'golden', 'tokenizer', 'world_with_tokenizer.pt'), map_location='cpu', weights_only=False)
    tk = fx['tokenizer_kwargs']
    mk = dict(fx['model_kwargs'], num_latent_tokens=tk['num_latent_tokens'])
    sd = {k: v for k, v in fx['state_dict'].items() if not k.startswith('video_tokenizer.')}
    tsd = {k[len('video_tokenizer.'):]: v for k, v in fx['state_dict'].items() if k.startswith('video_tokenizer.')}
    tcfg = TO.config_from_reference_kwargs(**tk)

    def obs_to_latents(world_model, obs, cache):
        tok_cache, t = cache if cache is not None else (None, 0)
        lat, tok_cache = TO.tokenize_step(tsd, tcfg, obs['image'], tok_cache, t)
        return lat[:, None], (tok_cache, t + 1)

    model = DynamicsWorldModel(**mk, precision='fp32')
    model.load_state_dict(sd, strict=True)
    before = {k: v.detach().clone() for k, v in model.state_dict().items()}
    trainer = SimTrainer(model, batch_size=2, epochs=1, learning_rate=1e-2)
    env = ToyImageEnv(batch=3, terminate_at=None)
    try:
        torch.manual_seed(0)
        episode = model.interact_with_env(env, env_is_vectorized=True, max_timesteps=2, obs_to_latents_fn=obs_to_latents).cpu()
        combined = combine_experiences([episode])
        # the first minibatch the trainer will draw, and its losses computed directly
        torch.manual_seed(7)
        first = next(iter(DataLoader(TensorDataset(torch.arange(combined.latents.shape[0])), batch_size=2, shuffle=True)))[0]
        pick = lambda t: t[first]
        batch = Experience(latents=pick(combined.latents), actions=Actions(pick(combined.actions.discrete), None),
                           log_probs=Actions(pick(combined.log_probs.discrete), None), agent_embed=pick(combined.agent_embed),
                           old_action_unembeds=Actions(pick(combined.old_action_unembeds.discrete), None), values=pick(combined.values),
                           rewards=pick(combined.rewards), step_size=combined.step_size, agent_index=combined.agent_index)
        pl, vl = model.learn_from_experience(batch)
        torch.manual_seed(7)
        losses = trainer.learn(combined)
    finally:
        model._release()
    assert len(losses) == 2 and int(trainer.step) == 2                      # 3 episodes in minibatches of 2, one epoch
    torch.testing.assert_close(losses[0][0], pl.detach(), atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(losses[0][1], vl.detach(), atol=1e-6, rtol=1e-5)
    moved = {k for k, v in model.state_dict().items() if not torch.equal(v, before[k])}
    assert moved and all(k.startswith(('policy_head.', 'value_head.')) or k == 'action_embedder.discrete_action_unembed' for k in moved), sorted(moved)


def test_interact_with_env_through_the_attached_tokenizer_on_the_simulator(on_simulator):
    """interact_with_env on image observations with the VideoTokenizer attached and no obs_to_latents_fn: each frame through the
    encoder's d4_tf_step over its time cache, then d4_observe - the REFERENCE's own episode (same CPU sampler draws) comes out."""
    from dreamer4_b200 import DynamicsWorldModel, VideoTokenizer
    from oracle.toy_env import ToyImageEnv
    fx = torch.load(os.path.join(HERE, 'golden', 'tokenizer', 'world_with_tokenizer.pt'), map_location='cpu', weights_only=False)
    ref = fx['interact'][2]                                   # single episode cut by max_timesteps: the bootstrap step runs too
    tok = VideoTokenizer(**fx['tokenizer_kwargs'], precision='fp32')
    model = DynamicsWorldModel(**fx['model_kwargs'], video_tokenizer=tok, precision='fp32')
    model.load_state_dict(fx['state_dict'], strict=True)
    try:
        torch.manual_seed(ref['seed'])
        exp = model.interact_with_env(ToyImageEnv(batch=None, terminate_at=ref['terminate_at']), max_timesteps=ref['max_timesteps'],
                                      env_is_vectorized=False)
    finally:
        model._release()
        tok._release()
    assert torch.equal(exp.actions.discrete, ref['actions']) and torch.equal(exp.lens, ref['lens'])
    assert torch.equal(exp.is_truncated, ref['is_truncated']) and torch.equal(exp.terminals, ref['terminals'])
    for name in ('latents', 'agent_embed', 'values', 'rewards'):
        torch.testing.assert_close(getattr(exp, name), ref[name], atol=5e-5, rtol=2e-4, msg=lambda m, n=name: f'{n}: {m}')


def test_world_model_with_tokenizer_on_the_simulator_reproduces_reference_golden(on_simulator):
    """The reference's own seeded run, drop-in: generate(prompt=video) -> decoded video of a DynamicsWorldModel with its tokenizer
    attached (reference dreamer4.py:6377-6387, 6694-6724: tokenize the prompt, prefill, roll out, decode).  On the CPU the host classes
    draw from torch's generator in the reference's order, so the same seed gives the reference's numbers."""
    from dreamer4_b200 import DynamicsWorldModel, VideoTokenizer
    fx = torch.load(os.path.join(HERE, 'golden', 'tokenizer', 'world_with_tokenizer.pt'), map_location='cpu', weights_only=False)
    tok = VideoTokenizer(**fx['tokenizer_kwargs'], precision='fp32')
    model = DynamicsWorldModel(**fx['model_kwargs'], video_tokenizer=tok, precision='fp32')
    model.load_state_dict(fx['state_dict'], strict=True)
    B = fx['prompt'].shape[0]
    try:
        ref = fx['prompted']
        torch.manual_seed(ref['seed'])
        video = model.generate(ref['time_steps'], batch_size=B, prompt=fx['prompt'])
        torch.testing.assert_close(video, ref['video'], atol=1e-4, rtol=2e-4)
    finally:
        model._release()
        tok._release()


def test_axial_space_time_transformer_on_the_simulator_reproduces_reference_golden(on_simulator, monkeypatch):
    """The stand-alone AxialSpaceTimeTransformer (exported by the reference package): state_dict layout, a 3-frame forward with its
    time-KV cache, and the same frames fed one at a time through `cache=` - against the reference module's own output."""
    from dreamer4_b200 import AxialSpaceTimeTransformer
    monkeypatch.setattr(AxialSpaceTimeTransformer, '_require_cuda', lambda self: None)
    monkeypatch.setattr(AxialSpaceTimeTransformer, '_stream', lambda self: C.c_void_p(0))
    fx = torch.load(os.path.join(HERE, 'golden', 'cache', 'axial_transformer.pt'), map_location='cpu', weights_only=False)
    m = AxialSpaceTimeTransformer(**fx['kwargs'], precision='fp32')
    assert set(m.state_dict()) == set(fx['state_dict'])
    m.load_state_dict(fx['state_dict'], strict=True)
    try:
        out, inter = m(fx['tokens'], return_intermediates=True)
        torch.testing.assert_close(out, fx['out'], atol=2e-5, rtol=1e-4)
        assert inter.token_count == fx['token_count']
        torch.testing.assert_close(inter.next_kv_cache, fx['kv'], atol=2e-5, rtol=1e-4)
        cache, frames = None, []
        for t in range(fx['tokens'].shape[1]):                      # the reference's incremental calling convention (:2957-2961)
            o, cache = m(fx['tokens'][:, :t + 1], cache=cache, return_intermediates=True)
            frames.append(o[:, -1])
        torch.testing.assert_close(torch.stack(frames, dim=1), fx['out'], atol=2e-5, rtol=1e-4)
    finally:
        m._release()


def test_video_tokenizer_with_64_latent_tokens_on_the_simulator_matches_oracle(on_simulator):
    """Frames of more than 64 tokens, 64 of them special - config 4's latent-token count (the golden fixtures have 3 and 6): the
    frame_attn.cu path with several key rounds per lane, the special-token mask over a 64-wide block, 64-row groups in the row maps."""
    from dreamer4_b200 import VideoTokenizer
    from oracle import tokenizer_oracle as TO
    kw = dict(dim=32, dim_latent=8, patch_size=4, image_size=16, num_latent_tokens=64, encoder_depth=2, decoder_depth=2, time_block_every=2,
              attn_heads=2, attn_dim_head=16)
    torch.manual_seed(3)
    tok = VideoTokenizer(**kw, precision='fp32')
    with torch.no_grad():
        for n, prm in tok.named_parameters():
            if n.endswith('gamma') or n.endswith('norm.weight'):
                prm.add_(torch.randn_like(prm) * 0.1)
            if n == 'latent_tokens':
                prm.mul_(30.)
    sd = {k: v.detach().clone() for k, v in tok.state_dict().items()}
    ocfg = TO.config_from_reference_kwargs(**kw)
    video, noise = torch.randn(1, 3, 2, 16, 16), torch.randn(1, 3, 2, 16, 16)
    want = TO.tokenize(sd, ocfg, video)
    try:
        torch.testing.assert_close(tok.tokenize(video), want, atol=2e-5, rtol=1e-4)
        torch.testing.assert_close(tok.decode(want, noise=noise), TO.decode(sd, ocfg, want, noise=noise), atol=5e-5, rtol=1e-4)
    finally:
        tok._release()
