"""Host-side bookkeeping of DynamicsWorldModel.generate and DynamicsWorldModelWrapper on CPU: the native calls are routed
to tests/fake_engine.py (the oracle's arithmetic behind the C-ABI's signatures), so every output must equal
oracle.generate's BIT FOR BIT - any difference is a host bug (prompt handling, action history, time-cache hand-back,
capacity growth, slicing), not rounding.  The CUDA kernels behind the same calls are covered by the -m gpu tests."""
import glob
import os

import pytest
import torch

from oracle import dreamer4_oracle as O
from fake_engine import install

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden', '*.pt')))
IDS = [os.path.basename(p)[:-3] for p in GOLDEN]


@pytest.fixture
def setup(monkeypatch):
    """setup(path) -> (fixture, product model on CPU, oracle config, fake engine).  The model's fake context is released
    before monkeypatch restores the real loader (a fake handle must never reach the real d4_ctx_destroy)."""
    from dreamer4_b200 import DynamicsWorldModel
    made = []

    def make(path):
        fx = torch.load(path, map_location='cpu', weights_only=False)
        model = DynamicsWorldModel(**fx['model_kwargs'], precision='fp32')
        model.load_state_dict(fx['state_dict'], strict=True)
        ocfg = O.config_from_reference_kwargs(**fx['model_kwargs'])
        made.append(model)
        return fx, model, ocfg, install(monkeypatch, model, ocfg)

    yield make
    for model in made:
        model._release()


def make_noise(cfg, T, B, seed):
    g = torch.Generator().manual_seed(seed)
    A = sum(cfg.num_discrete_actions)
    return dict(latent=torch.randn(T, B, cfg.num_latent_tokens, cfg.dim_latent, generator=g),
                action_uniform=torch.rand(T, B, max(A, 1), generator=g)[..., :A],
                terminal_uniform=torch.rand(T, B, generator=g))


def injected(noise):
    return O.InjectedNoise(noise['latent'], noise['action_uniform'], noise['terminal_uniform'])


def same(exp, ref, tc=None):
    assert torch.equal(exp.latents, ref.latents)
    assert torch.equal(exp.rewards, ref.rewards)
    assert torch.equal(exp.lens, ref.lens) and torch.equal(exp.terminals, ref.terminals) and torch.equal(exp.is_truncated, ref.is_truncated)
    assert torch.equal(exp.episode_return, ref.episode_return)
    if ref.agent_embed is None:
        assert exp.agent_embed is None or exp.agent_embed.shape[1] == 0
    else:
        assert torch.equal(exp.agent_embed, ref.agent_embed)
    if ref.actions is None:
        assert exp.actions is None
    else:
        assert torch.equal(exp.actions.discrete, ref.actions)
        assert torch.equal(exp.log_probs.discrete, ref.log_probs) and torch.equal(exp.values, ref.values)
        # the oracle unembeds all frames in one matmul (dreamer4.py:6749-6750), the engine frame by frame: same values up to blocking
        torch.testing.assert_close(exp.old_action_unembeds.discrete, ref.old_action_unembeds, atol=1e-5, rtol=1e-5)
    if tc is not None:
        kv = torch.stack([torch.stack(layer) for layer in ref.kv_cache])
        assert tc.main.token_count == ref.latents.shape[1]
        assert torch.equal(tc.main.next_kv_cache, kv)


class ProductOrderNoise(O.TorchRNGNoise):
    """torch's global generator in the product's draw order: the terminal draw is rand < p (the oracle's default mirrors
    the reference's torch.bernoulli, which consumes the CPU generator differently)."""

    def terminal(self, frame, probs):
        return torch.rand(probs.shape) < probs


FLAGS = dict(return_rewards_per_frame=True, return_agent_actions=True, return_log_probs_and_values=True, return_time_cache=True)


@pytest.mark.parametrize('path', GOLDEN, ids=IDS)
def test_plain_rollout(path, setup):
    fx, model, ocfg, _ = setup(path)
    gk = dict(fx['gen_kwargs'])
    T, B = gk.pop('time_steps'), gk.pop('batch_size')
    noise = make_noise(ocfg, T, B, 1)
    ref = O.generate(fx['state_dict'], ocfg, T, B, noise=injected(noise), **gk)
    exp, tc = model.generate(T, batch_size=B, noise=noise, **FLAGS, **gk)
    same(exp, ref, tc)


@pytest.mark.parametrize('keep_view', [True, False], ids=['in_place', 'cloned_cache'])
@pytest.mark.parametrize('path', GOLDEN, ids=IDS)
def test_resumed_rollout(path, keep_view, setup):
    """generate(P) then generate(T, prompt..., time_cache) == one oracle rollout resumed the same way; the cache is taken
    either as the live view of the KV buffer (no copy) or as a clone (copied in)."""
    fx, model, ocfg, fake = setup(path)
    B, T, P = fx['gen_kwargs']['batch_size'], 5, 2
    term = dict(return_terminals=True) if fx['gen_kwargs'].get('return_terminals') else {}
    noise = make_noise(ocfg, T, B, 2)
    head_ref = O.generate(fx['state_dict'], ocfg, P, B, noise=injected(noise))
    ref = O.generate(fx['state_dict'], ocfg, T, B, noise=injected(noise), prompt_latents=head_ref.latents, prompt_actions=head_ref.actions,
                     prompt_rewards=head_ref.rewards, kv_cache=head_ref.kv_cache, **term)
    head, tc = model.generate(P, batch_size=B, noise=noise, **FLAGS)
    same(head, head_ref, tc)
    if not keep_view:
        tc = type(tc)(main=type(tc.main)(next_kv_cache=tc.main.next_kv_cache.clone(), token_count=tc.main.token_count))
    exp, tc2 = model.generate(T, batch_size=B, noise=noise, prompt_latents=head.latents, prompt_discrete_actions=head.actions.discrete,
                              prompt_rewards=head.rewards, time_cache=tc, **FLAGS, **term)
    same(exp, ref, tc2)
    assert fake.calls['pass_'] == 0                      # resumed: no prefill
    assert exp.agent_embed.shape[1] == exp.latents.shape[1] - P and exp.actions.discrete.shape[1] == exp.latents.shape[1]


@pytest.mark.parametrize('n_prompt_actions', [2, 1, 0], ids=['actions_P', 'actions_P-1', 'no_actions'])
@pytest.mark.parametrize('path', [p for p in GOLDEN if 'terminals' not in p], ids=[i for i in IDS if 'terminals' not in i])
def test_cold_prompt(path, n_prompt_actions, setup):
    """No cache: one clean d4_pass per prompt frame rebuilds it; the action history follows the reference's right-pad /
    shift rule for every prompt-action count (dreamer4.py:6519-6522, 7111-7126)."""
    fx, model, ocfg, fake = setup(path)
    B, T, P = fx['gen_kwargs']['batch_size'], 4, 2
    noise = make_noise(ocfg, T, B, 3)
    head = O.generate(fx['state_dict'], ocfg, P, B, noise=injected(noise))
    pa = head.actions[:, :n_prompt_actions] if n_prompt_actions > 0 else None
    ref = O.generate(fx['state_dict'], ocfg, T, B, noise=injected(noise), prompt_latents=head.latents, prompt_actions=pa, prompt_rewards=head.rewards)
    exp, tc = model.generate(T, batch_size=B, noise=noise, prompt_latents=head.latents, prompt_discrete_actions=pa, prompt_rewards=head.rewards, **FLAGS)
    same(exp, ref, tc)
    assert fake.calls['pass_'] == P and fake.calls['frame'] == T - P


@pytest.mark.parametrize('path', GOLDEN, ids=IDS)
def test_env_wrapper_steps(path, setup):
    """reset() + step(action) x n == the oracle resumed one frame at a time with the same supplied actions; the context is
    created once for the 1-frame reset and once more (64-frame capacity) for all the steps."""
    from dreamer4_b200 import DynamicsWorldModelWrapper
    fx, model, ocfg, fake = setup(path)
    B, steps = 2, 4
    sizes = ocfg.num_discrete_actions
    g = torch.Generator().manual_seed(5)
    supplied = torch.stack([torch.randint(0, n, (steps, B), generator=g) for n in sizes], dim=-1)       # (steps, B, na)
    env = DynamicsWorldModelWrapper(model, num_generation_steps=4)

    torch.manual_seed(11)
    obs, info = env.reset(batch_size=B)
    outs = [env.step(supplied[i]) for i in range(steps)]

    torch.manual_seed(11)
    ref = O.generate(fx['state_dict'], ocfg, 1, B, noise=ProductOrderNoise(), return_terminals=True, return_agent_actions=False)
    assert torch.equal(obs, ref.latents[:, -1])
    for i in range(steps):
        ref = O.generate(fx['state_dict'], ocfg, i + 2, B, noise=ProductOrderNoise(), return_terminals=True, return_agent_actions=False, prompt_latents=ref.latents,
                         prompt_actions=supplied[:i + 1].transpose(0, 1), prompt_rewards=ref.rewards, kv_cache=ref.kv_cache)
        obs, reward, terminated, truncated, info = outs[i]
        assert torch.equal(obs, ref.latents[:, -1]) and torch.equal(reward, ref.rewards[:, -1])
        assert torch.equal(terminated, ref.terminals) and not truncated.any()
        assert info['experience'].actions is None
    same(outs[-1][-1]['experience'], ref, env._time_cache)
    assert fake.calls['ctx_create'] == 2 and fake.calls['pass_'] == 0


def test_stale_view_is_rejected_and_capacity_grows(setup):
    fx, model, ocfg, fake = setup(GOLDEN[0])
    B = 2
    args = lambda e: dict(prompt_latents=e.latents, prompt_discrete_actions=e.actions.discrete, prompt_rewards=e.rewards)
    torch.manual_seed(0)
    a, tc_a = model.generate(2, batch_size=B, **FLAGS)
    b, tc_b = model.generate(3, batch_size=B, time_cache=tc_a, **args(a), **FLAGS)            # grows to a 64-frame buffer, copies
    assert model._ctx_key[1] == 64 and fake.calls['ctx_create'] == 2
    c, tc_c = model.generate(4, batch_size=B, time_cache=tc_b, **args(b), **FLAGS)            # in place
    assert fake.calls['ctx_create'] == 2
    with pytest.raises(ValueError, match='stale'):                                            # tc_b's frames were extended since
        model.generate(5, batch_size=B, time_cache=tc_b, **args(b), **FLAGS)
    with pytest.raises(AssertionError):                                                       # cache / prompt length mismatch
        model.generate(5, batch_size=B, time_cache=tc_c, **args(b), **FLAGS)
    with pytest.raises(AssertionError):                                                       # nothing to generate
        model.generate(4, batch_size=B, time_cache=tc_c, **args(c), **FLAGS)


WORLD = os.path.join(os.path.dirname(__file__), 'golden', 'tokenizer', 'world_with_tokenizer.pt')


@pytest.mark.parametrize('case', range(4), ids=['vec_mixed_bootstrap', 'vec_all_terminated', 'single_truncated', 'single_terminated'])
def test_interact_with_env(case, monkeypatch):
    """interact_with_env's bookkeeping (termination / truncation / bootstrap padding, previous-action conditioning, record
    slicing, env action formatting) against oracle.interact_with_env on the deterministic toy env, observations tokenized by
    the oracle's incremental tokenizer through `obs_to_latents_fn`.  One d4_observe per env step (+1 for the bootstrap)."""
    from dreamer4_b200 import DynamicsWorldModel
    from oracle import tokenizer_oracle as TO
    from oracle.toy_env import ToyImageEnv
    fx = torch.load(WORLD, map_location='cpu', weights_only=False)
    ref_case = fx['interact'][case]
    vectorized, terminate_at, max_timesteps = ref_case['vectorized'], ref_case['terminate_at'], ref_case['max_timesteps']
    tk = fx['tokenizer_kwargs']
    mk = dict(fx['model_kwargs'], num_latent_tokens=tk['num_latent_tokens'])
    sd = {k: v for k, v in fx['state_dict'].items() if not k.startswith('video_tokenizer.')}
    tsd = {k[len('video_tokenizer.'):]: v for k, v in fx['state_dict'].items() if k.startswith('video_tokenizer.')}
    ocfg, tcfg = O.config_from_reference_kwargs(**mk), TO.config_from_reference_kwargs(**tk)
    model = DynamicsWorldModel(**mk, precision='fp32')
    model.load_state_dict(sd, strict=True)
    fake = install(monkeypatch, model, ocfg)

    def obs_to_latents(world_model, obs, cache):
        assert world_model is model
        frame = obs['image'] if vectorized else obs['image'][None]
        tok_cache, t = cache if cache is not None else (None, 0)
        lat, tok_cache = TO.tokenize_step(tsd, tcfg, frame, tok_cache, t)
        return lat[:, None], (tok_cache, t + 1)

    try:
        torch.manual_seed(ref_case['seed'])
        exp = model.interact_with_env(ToyImageEnv(batch=3 if vectorized else None, terminate_at=terminate_at), max_timesteps=max_timesteps,
                                      env_is_vectorized=vectorized, obs_to_latents_fn=obs_to_latents)
        torch.manual_seed(ref_case['seed'])
        ref = O.interact_with_env(fx['state_dict'], ocfg, (tsd, tcfg), ToyImageEnv(batch=3 if vectorized else None, terminate_at=terminate_at),
                                  max_timesteps=max_timesteps, env_is_vectorized=vectorized)
        for name in ('latents', 'agent_embed', 'rewards', 'values', 'episode_return', 'lens', 'terminals', 'is_truncated'):
            assert torch.equal(getattr(exp, name), getattr(ref, name)), name
        assert torch.equal(exp.actions.discrete, ref.actions) and torch.equal(exp.log_probs.discrete, ref.log_probs)
        torch.testing.assert_close(exp.old_action_unembeds.discrete, ref.old_action_unembeds, atol=1e-5, rtol=1e-5)
        assert not exp.is_from_world_model and exp.video.shape[2] == exp.rewards.shape[1]
        # ... and, transitively, the reference's own episode (the oracle is pinned to it in test_tokenizer_oracle_golden.py)
        assert torch.equal(exp.actions.discrete, ref_case['actions']) and torch.equal(exp.lens, ref_case['lens'])
        torch.testing.assert_close(exp.values, ref_case['values'], atol=2e-5, rtol=1e-4)
        assert fake.calls['observe'] == exp.latents.shape[1] and fake.calls['pass_'] == 0
    finally:
        model._release()


# ------------------------------------------------------------------------------------------------ attached video tokenizer
# DynamicsWorldModel(video_tokenizer=VideoTokenizer(...)): state_dict layout with the tokenizer as a submodule, generate(prompt=
# video) and return_decoded_video (reference dreamer4.py:6377-6387, 6694-6724).  The tokenizer's two entry points are replaced by
# the oracle's here (a fake tokenizer engine, like fake_engine.py for the dynamics model): what is under test is the wiring.

WORLD = os.path.join(os.path.dirname(__file__), 'golden', 'tokenizer', 'world_with_tokenizer.pt')


@pytest.fixture
def world(monkeypatch):
    from dreamer4_b200 import DynamicsWorldModel, VideoTokenizer
    from oracle import tokenizer_oracle as TO
    fx = torch.load(WORLD, map_location='cpu', weights_only=False)
    tok = VideoTokenizer(**fx['tokenizer_kwargs'])
    model = DynamicsWorldModel(fx['model_kwargs']['dim'], fx['model_kwargs']['dim_latent'], tok, precision='fp32',       # third positional, as the reference (:4666)
                               **{k: v for k, v in fx['model_kwargs'].items() if k not in ('dim', 'dim_latent')})
    assert model.cfg.num_latent_tokens == tok.num_latent_tokens                  # taken from the tokenizer (reference :4801)
    assert model.video_tokenizer is not tok and not any(p.requires_grad for p in model.video_tokenizer.parameters())      # copy_video_tokenizer (:4789-4792)
    tok = model.video_tokenizer
    model.load_state_dict(fx['state_dict'], strict=True)                         # video_tokenizer.* keys included
    tsd = {k[len('video_tokenizer.'):]: v for k, v in fx['state_dict'].items() if k.startswith('video_tokenizer.')}
    tcfg = TO.config_from_reference_kwargs(**fx['tokenizer_kwargs'])
    monkeypatch.setattr(tok, 'tokenize', lambda video, **kw: TO.tokenize(tsd, tcfg, video))
    dec_noise = {}                                                               # test-provided start noise of the flow decoder
    monkeypatch.setattr(tok, 'decode', lambda latents, height=None, width=None, **kw: TO.decode(tsd, tcfg, latents, noise=dec_noise['x']))
    ocfg = O.config_from_reference_kwargs(num_latent_tokens=tok.num_latent_tokens, **fx['model_kwargs'])
    fake = install(monkeypatch, model, ocfg)
    yield fx, model, ocfg, (tsd, tcfg), dec_noise
    model._release()


def test_video_prompt_and_decoded_video(world):
    fx, model, ocfg, tokenizer, dec_noise = world
    sd = fx['state_dict']
    B, T = fx['prompt'].shape[0], fx['prompted']['time_steps']
    noise = make_noise(model.cfg, T, B, seed=11)
    tcfg = tokenizer[1]
    dec_noise['x'] = torch.randn(B, tcfg.channels, T, tcfg.image_height, tcfg.image_width)        # the decoder's randn (reference :4204)
    ref = O.generate(sd, ocfg, T, B, noise=O.InjectedNoise(noise['latent'], noise['action_uniform'], noise['terminal_uniform'], dec_noise['x']),
                     tokenizer=tokenizer, prompt=fx['prompt'], return_agent_actions=False, return_decoded_video=True)
    video = model.generate(T, batch_size=B, prompt=fx['prompt'], noise=noise)    # tokenizer attached: the decoded video comes back (:6694, 6721)
    assert torch.is_tensor(video) and video.shape == ref.video.shape
    assert torch.equal(video, ref.video)
    latents = model.generate(T, batch_size=B, prompt=fx['prompt'], noise=noise, return_decoded_video=False)
    assert torch.equal(latents, ref.latents)


def test_dream_returns_experience_with_video(world):
    fx, model, ocfg, tokenizer, dec_noise = world
    sd = fx['state_dict']
    B, T = 2, 4
    noise = make_noise(model.cfg, T, B, seed=12)
    tcfg = tokenizer[1]
    dec_noise['x'] = torch.randn(B, tcfg.channels, T, tcfg.image_height, tcfg.image_width)
    ref = O.generate(sd, ocfg, T, B, noise=O.InjectedNoise(noise['latent'], noise['action_uniform'], noise['terminal_uniform'], dec_noise['x']),
                     tokenizer=tokenizer, return_decoded_video=True)
    exp = model.generate(T, batch_size=B, noise=noise, return_rewards_per_frame=True, return_agent_actions=True, return_log_probs_and_values=True)
    same(exp, ref)
    assert torch.equal(exp.video, ref.video)


@pytest.mark.parametrize('case', [0, 2], ids=['vec_mixed_bootstrap', 'single_truncated'])
def test_interact_with_env_through_the_attached_tokenizer(case, world, monkeypatch):
    """Without obs_to_latents_fn, image observations go through the attached tokenizer one frame per env step over its time cache
    (reference dreamer4.py:5588) - same episode as the reference's."""
    from oracle import tokenizer_oracle as TO
    from oracle.toy_env import ToyImageEnv
    fx, model, ocfg, (tsd, tcfg), _ = world
    ref_case = fx['interact'][case]
    vectorized = ref_case['vectorized']

    def tokenize(video, time_cache=None, return_time_cache=False):          # the oracle's incremental tokenizer behind the product's signature
        assert video.ndim == 5 and video.shape[2] == 1 and return_time_cache
        tok_cache, t = time_cache if time_cache is not None else (None, 0)
        lat, tok_cache = TO.tokenize_step(tsd, tcfg, video[:, :, 0], tok_cache, t)
        return lat[:, None], (tok_cache, t + 1)

    monkeypatch.setattr(model.video_tokenizer, 'tokenize', tokenize)
    torch.manual_seed(ref_case['seed'])
    exp = model.interact_with_env(ToyImageEnv(batch=3 if vectorized else None, terminate_at=ref_case['terminate_at']),
                                  max_timesteps=ref_case['max_timesteps'], env_is_vectorized=vectorized)
    assert torch.equal(exp.actions.discrete, ref_case['actions']) and torch.equal(exp.lens, ref_case['lens'])
    assert torch.equal(exp.is_truncated, ref_case['is_truncated']) and torch.equal(exp.terminals, ref_case['terminals'])
    torch.testing.assert_close(exp.latents, ref_case['latents'], atol=2e-5, rtol=1e-4)
    torch.testing.assert_close(exp.values, ref_case['values'], atol=2e-5, rtol=1e-4)


def test_env_wrapper_returns_decoded_frames_with_a_tokenizer(world, monkeypatch):
    """DynamicsWorldModelWrapper over a model with a tokenizer: observations are the newest DECODED frame (reference env.py:441, 505),
    a video prompt seeds the episode."""
    from dreamer4_b200 import DynamicsWorldModelWrapper
    from oracle import tokenizer_oracle as TO
    fx, model, ocfg, (tsd, tcfg), _ = world
    decode = lambda latents, height=None, width=None, **kw: TO.decode(tsd, tcfg, latents, noise=torch.zeros(latents.shape[0], tcfg.channels, latents.shape[1], tcfg.image_height, tcfg.image_width))
    monkeypatch.setattr(model.video_tokenizer, 'decode', decode)
    env = DynamicsWorldModelWrapper(model, num_generation_steps=4)
    assert env.image_size == tcfg.image_height
    obs, _ = env.reset(batch_size=2, seed=1)
    assert obs.shape == (2, tcfg.channels, tcfg.image_height, tcfg.image_width)
    obs, reward, terminated, truncated, info = env.step(torch.tensor([1, 2]))
    gen = info['experience']
    assert gen.latents.shape[1] == 2 and torch.equal(obs, decode(gen.latents)[:, :, -1]) and reward.shape == (2,)
    prompted = DynamicsWorldModelWrapper(model, prompt=fx['prompt'], num_generation_steps=4)
    obs, _ = prompted.reset()
    P = fx['prompt'].shape[2]
    assert prompted._latents.shape[:2] == (fx['prompt'].shape[0], P + 1)
    torch.testing.assert_close(prompted._latents[:, :P], TO.tokenize(tsd, tcfg, fx['prompt']).clamp(-1, 1))
    obs2, *_ = prompted.step(torch.zeros(fx['prompt'].shape[0], dtype=torch.long))
    assert prompted._latents.shape[1] == P + 2 and obs2.shape == obs.shape


def test_cache_continuation_without_prompt(monkeypatch):
    """generate(time_cache=...) without prompt latents (the reference's tests/test_dreamer.py::test_cache_generate): each call returns
    time_steps NEW frames imagined on top of the cached ones.  Three chained calls against the oracle on the same draws, and - since
    the oracle is pinned to them - the reference's own chained calls (tests/golden/cache/cache_continue.pt)."""
    from dreamer4_b200 import DynamicsWorldModel
    fx = torch.load(os.path.join(os.path.dirname(__file__), 'golden', 'cache', 'cache_continue.pt'), map_location='cpu', weights_only=False)
    model = DynamicsWorldModel(**fx['model_kwargs'], precision='fp32')
    model.load_state_dict(fx['state_dict'], strict=True)
    ocfg = O.config_from_reference_kwargs(**fx['model_kwargs'])
    install(monkeypatch, model, ocfg)
    try:
        tc, ocache = None, None
        for i, call in enumerate(fx['calls']):
            T = call['time_steps']
            noise = make_noise(model.cfg, T, 2, seed=40 + i)
            ref = O.generate(fx['state_dict'], ocfg, T, 2, noise=injected(noise), kv_cache=ocache)
            ocache = ref.kv_cache
            exp, tc = model.generate(T, batch_size=2, noise=noise, time_cache=tc, return_time_cache=True, return_rewards_per_frame=True,
                                     return_agent_actions=True, return_log_probs_and_values=True)
            assert exp.latents.shape[1] == T and tc.main.token_count == call['token_count']
            same(exp, ref)
            kv = torch.stack([torch.stack(layer) for layer in ref.kv_cache])
            assert torch.equal(tc.main.next_kv_cache, kv)
    finally:
        model._release()
