"""GPU parity of prompted / resumed rollouts and of the env wrapper built on them (reference dreamer4.py:6377-6402,
env.py:353-553): generate(prompt_latents, prompt_discrete_actions, prompt_rewards[, time_cache]) through the C-ABI against
(a) what the reference itself produced for the same continuation (tests/golden, its CPU RNG stream replayed) and (b) the
oracle on injected noise.  Tolerances as in tests/test_gpu_parity.py: exact-fp32 mode 5e-5 + 2e-4 rel, sampled actions bit-exact."""
import glob
import os

import pytest
import torch

from oracle import dreamer4_oracle as O

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden', '*.pt')))
IDS = [os.path.basename(p)[:-3] for p in GOLDEN]
TOL = dict(atol=5e-5, rtol=2e-4)
FLAGS = dict(return_rewards_per_frame=True, return_agent_actions=True, return_log_probs_and_values=True, return_time_cache=True)


def load(path):
    return torch.load(path, map_location='cpu', weights_only=False)


PROMPTED = [p for p in GOLDEN if 'prompted' in load(p)['out']]


def build_model(fx, **extra):
    from dreamer4_b200 import DynamicsWorldModel
    extra.setdefault('precision', 'fp32')
    model = DynamicsWorldModel(**fx['model_kwargs'], **extra)
    model.load_state_dict(fx['state_dict'], strict=True)
    return model.cuda()


def make_noise(cfg, T, B, seed):
    g = torch.Generator().manual_seed(seed)
    A = sum(cfg.num_discrete_actions)
    return dict(latent=torch.randn(T, B, cfg.num_latent_tokens, cfg.dim_latent, generator=g),
                action_uniform=torch.rand(T, B, max(A, 1), generator=g)[..., :A],
                terminal_uniform=torch.rand(T, B, generator=g))


def to_cuda(noise):
    return {k: v.cuda() for k, v in noise.items()}


def injected(noise):
    return O.InjectedNoise(noise['latent'], noise['action_uniform'], noise['terminal_uniform'])


def cache_of(kv, frames):
    from dreamer4_b200.experience import DynamicsIntermediates, TransformerIntermediates
    return DynamicsIntermediates(main=TransformerIntermediates(next_kv_cache=kv.cuda(), token_count=frames))


def close(exp, ref, tc, P, heads=True, tol=TOL):
    """exp: product Experience (cuda); ref: OracleExperience or a golden dict with the same field names."""
    get = (lambda k: ref[k]) if isinstance(ref, dict) else (lambda k: getattr(ref, k))
    T = get('latents').shape[1]
    assert exp.latents.shape == get('latents').shape
    for name in ('latents', 'agent_embed', 'rewards', 'episode_return'):
        torch.testing.assert_close(getattr(exp, name).cpu(), get(name), **tol, msg=lambda m, n=name: f'{n}: {m}')
    assert exp.agent_embed.shape[1] == T - P                       # new frames only
    assert torch.equal(exp.lens.cpu(), get('lens')) and torch.equal(exp.terminals.cpu(), get('terminals'))
    if heads:
        assert torch.equal(exp.actions.discrete.cpu(), get('actions'))             # prompt + sampled, bit exact
        torch.testing.assert_close(exp.log_probs.discrete.cpu(), get('log_probs'), **tol)
        torch.testing.assert_close(exp.values.cpu(), get('values'), **tol)
        torch.testing.assert_close(exp.old_action_unembeds.discrete.cpu(), get('old_action_unembeds'), **tol)
    else:
        assert exp.actions is None and exp.log_probs is None and exp.values is None
    kv = get('kv_cache')
    kv = kv if torch.is_tensor(kv) else torch.stack([torch.stack(layer) for layer in kv])
    assert tc.main.token_count == T
    torch.testing.assert_close(tc.main.next_kv_cache.cpu(), kv, **tol)


@pytest.mark.parametrize('flow', ['resume', 'cold', 'env_step'])
@pytest.mark.parametrize('path', PROMPTED, ids=[os.path.basename(p)[:-3] for p in PROMPTED])
def test_prompted_matches_reference_golden(path, flow):
    """The reference's own continuation of its own 2-frame rollout: resumed over its time cache, rebuilt cold from the
    prompt (its uncached multi-frame forward vs this path's per-frame prefill), and env.py's single step with a supplied
    action.  Its CPU RNG stream is replayed (randn latent, rand per action type when the policy acts, randn context)."""
    fx = load(path)
    model = build_model(fx)
    cfg = model.cfg
    pr = fx['out']['prompted']
    head, ref, P = pr['head'], pr[flow], pr['P']
    B, T = ref['latents'].shape[:2]
    heads = flow != 'env_step'
    torch.manual_seed(ref['seed'])
    lat = torch.zeros(T, B, cfg.num_latent_tokens, cfg.dim_latent)
    au = torch.zeros(T, B, sum(cfg.num_discrete_actions))
    for t in range(P, T):
        lat[t] = torch.randn(B, 1, 1, cfg.num_latent_tokens, cfg.dim_latent).reshape(B, cfg.num_latent_tokens, cfg.dim_latent)
        if heads:
            au[t] = torch.cat([torch.rand(B, 1, n).reshape(B, n) for n in cfg.num_discrete_actions], dim=-1)
        torch.randn(B, 1, 1, cfg.num_latent_tokens, cfg.dim_latent)
    noise = dict(latent=lat, action_uniform=au, terminal_uniform=torch.zeros(T, B))
    kw = FLAGS if heads else dict(return_rewards_per_frame=True, return_time_cache=True)
    exp, tc = model.generate(T, batch_size=B, noise=to_cuda(noise), prompt_latents=head['latents'].cuda(),
                             prompt_discrete_actions=head['actions'].cuda(), prompt_rewards=head['rewards'].cuda(),
                             time_cache=None if flow == 'cold' else cache_of(head['kv_cache'], P), **kw)
    close(exp, ref, tc, P, heads=heads)
    assert torch.equal(exp.latents[:, :P].cpu(), head['latents'])


@pytest.mark.parametrize('precision', ['fp32', 'tf32x3'])
def test_prompted_midsize_matches_oracle(precision):
    """Resumed (in-place view of the live KV buffer) and cold continuations on a model wide enough for the tcgen05 path."""
    from test_gpu_parity import MID, _mid_model
    model, sd = _mid_model(precision)
    ocfg = O.config_from_reference_kwargs(**MID)
    T, P, B = 5, 2, 6
    noise = make_noise(model.cfg, T, B, seed=21)
    head_ref = O.generate(sd, ocfg, P, B, noise=injected(noise))
    prompt = dict(prompt_latents=head_ref.latents, prompt_actions=head_ref.actions, prompt_rewards=head_ref.rewards)
    resumed_ref = O.generate(sd, ocfg, T, B, noise=injected(noise), kv_cache=head_ref.kv_cache, **prompt)
    cold_ref = O.generate(sd, ocfg, T, B, noise=injected(noise), **prompt)
    tol = TOL if precision == 'fp32' else dict(atol=2e-4, rtol=2e-4)

    head, tc = model.generate(P, batch_size=B, noise=to_cuda(noise), **FLAGS)
    args = dict(prompt_latents=head.latents, prompt_discrete_actions=head.actions.discrete, prompt_rewards=head.rewards)
    exp, tc2 = model.generate(T, batch_size=B, noise=to_cuda(noise), time_cache=tc, **args, **FLAGS)
    close(exp, resumed_ref, tc2, P, tol=tol)
    exp, tc3 = model.generate(T, batch_size=B, noise=to_cuda(noise), **args, **FLAGS)
    close(exp, cold_ref, tc3, P, tol=tol)


def test_resume_across_capacity_growth_equals_one_shot():
    """3 frames, then resumed to 70: the 3-frame KV buffer is replaced by a 128-frame one (copy), and the result equals the
    one-shot 70-frame rollout on the same noise."""
    fx = load(GOLDEN[0])
    model = build_model(fx)
    T, P, B = 70, 3, 2
    noise = to_cuda(make_noise(model.cfg, T, B, seed=31))
    one, tc_one = model.generate(T, batch_size=B, noise=noise, **FLAGS)
    one_kv = tc_one.main.next_kv_cache.clone()
    model._release()                     # a context whose capacity already covers the call is reused: start over so that the 3-frame one is built
    head, tc = model.generate(P, batch_size=B, noise=noise, **FLAGS)
    assert model._ctx_key[1] == P
    two, tc_two = model.generate(T, batch_size=B, noise=noise, time_cache=tc, prompt_latents=head.latents,
                                 prompt_discrete_actions=head.actions.discrete, prompt_rewards=head.rewards, **FLAGS)
    assert model._ctx_key[1] == 128
    tight = dict(atol=1e-6, rtol=1e-6)
    assert torch.equal(two.actions.discrete, one.actions.discrete)
    torch.testing.assert_close(two.latents, one.latents, **tight)
    torch.testing.assert_close(two.rewards, one.rewards, **tight)
    torch.testing.assert_close(two.agent_embed, one.agent_embed[:, P:], **tight)
    torch.testing.assert_close(two.values, one.values[:, P:], **tight)
    torch.testing.assert_close(tc_two.main.next_kv_cache, one_kv, **tight)


@pytest.mark.parametrize('path', GOLDEN, ids=IDS)
def test_env_wrapper_matches_oracle(path):
    """DynamicsWorldModelWrapper.reset + 4 steps with supplied actions against the oracle resumed frame by frame on the very
    draws the wrapper consumed (the CUDA generator replayed in generate's order: randn latent, rand terminal, randn context)."""
    from dreamer4_b200 import DynamicsWorldModelWrapper
    fx = load(path)
    model = build_model(fx)
    cfg = model.cfg
    ocfg = O.config_from_reference_kwargs(**fx['model_kwargs'])
    B, steps = 3, 4
    g = torch.Generator().manual_seed(5)
    supplied = torch.stack([torch.randint(0, n, (steps, B), generator=g) for n in cfg.num_discrete_actions], dim=-1)     # (steps, B, na)
    env = DynamicsWorldModelWrapper(model, num_generation_steps=4)
    torch.manual_seed(41)
    obs0, _ = env.reset(batch_size=B)
    outs = [env.step(supplied[i].cuda()) for i in range(steps)]

    torch.manual_seed(41)
    lat, tu = [], []
    for _ in range(steps + 1):
        lat.append(torch.randn(B, cfg.num_latent_tokens, cfg.dim_latent, device='cuda'))
        tu.append(torch.rand(B, device='cuda') if model.predict_terminals else torch.zeros(B, device='cuda'))
        torch.randn(B, cfg.num_latent_tokens, cfg.dim_latent, device='cuda')
    noise = O.InjectedNoise(torch.stack(lat).cpu(), None, torch.stack(tu).cpu())

    ref = O.generate(fx['state_dict'], ocfg, 1, B, noise=noise, return_terminals=True, return_agent_actions=False)
    torch.testing.assert_close(obs0.cpu(), ref.latents[:, -1], **TOL)
    for i in range(steps):
        ref = O.generate(fx['state_dict'], ocfg, i + 2, B, noise=noise, return_terminals=True, return_agent_actions=False,
                         prompt_latents=ref.latents, prompt_actions=supplied[:i + 1].transpose(0, 1), prompt_rewards=ref.rewards, kv_cache=ref.kv_cache)
        obs, reward, terminated, truncated, info = outs[i]
        torch.testing.assert_close(obs.cpu(), ref.latents[:, -1], **TOL)
        torch.testing.assert_close(reward.cpu(), ref.rewards[:, -1], **TOL)
        assert torch.equal(terminated.cpu(), ref.terminals) and not truncated.any()
    close(outs[-1][-1]['experience'], ref, env._time_cache, P=steps, heads=False)


def test_stale_time_cache_is_rejected():
    fx = load(GOLDEN[0])
    model = build_model(fx)
    B = 2
    args = lambda e: dict(prompt_latents=e.latents, prompt_discrete_actions=e.actions.discrete, prompt_rewards=e.rewards)
    a, tc_a = model.generate(2, batch_size=B, **FLAGS)
    b, tc_b = model.generate(3, batch_size=B, time_cache=tc_a, **args(a), **FLAGS)
    model.generate(4, batch_size=B, time_cache=tc_b, **args(b), **FLAGS)
    with pytest.raises(ValueError, match='stale'):
        model.generate(5, batch_size=B, time_cache=tc_b, **args(b), **FLAGS)
