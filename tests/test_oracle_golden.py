"""The oracle (oracle/dreamer4_oracle.py) against golden vectors produced by executing the
reference's own dreamer4.py (oracle/make_golden.py).  CPU only."""
import glob
import os

import pytest
import torch

from oracle import dreamer4_oracle as O

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden', '*.pt')))


def load(path):
    return torch.load(path, map_location='cpu', weights_only=False)


@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p)[:-3] for p in GOLDEN])
def test_generate_matches_reference(path):
    fx = load(path)
    cfg = O.config_from_reference_kwargs(**fx['model_kwargs'])
    gk = dict(fx['gen_kwargs'])
    torch.manual_seed(fx['gen_seed'])
    exp = O.generate(fx['state_dict'], cfg, gk.pop('time_steps'), gk.pop('batch_size'), **gk)
    ref = fx['out']
    assert exp.latents.shape == ref['latents'].shape
    assert torch.equal(exp.actions, ref['actions'])                      # sampled indices: bit exact
    assert torch.equal(exp.lens, ref['lens'])
    assert torch.equal(exp.terminals, ref['terminals'])
    assert torch.equal(exp.is_truncated, ref['is_truncated'])
    assert exp.step_size == ref['step_size']
    tol = dict(atol=2e-5, rtol=1e-4)
    torch.testing.assert_close(exp.latents, ref['latents'], **tol)
    torch.testing.assert_close(exp.agent_embed, ref['agent_embed'], **tol)
    torch.testing.assert_close(exp.rewards, ref['rewards'], **tol)
    torch.testing.assert_close(exp.values, ref['values'], **tol)
    torch.testing.assert_close(exp.log_probs, ref['log_probs'], **tol)
    torch.testing.assert_close(exp.old_action_unembeds, ref['old_action_unembeds'], **tol)
    torch.testing.assert_close(exp.episode_return, ref['episode_return'], **tol)
    # KV cache: reference layout (y, 2, b*S, h, T, d)  (D4:3255-3265)
    kv = torch.stack([torch.stack(layer) for layer in exp.kv_cache])
    assert kv.shape == ref['kv_cache'].shape
    torch.testing.assert_close(kv, ref['kv_cache'], **tol)
    assert ref['token_count'] == exp.latents.shape[1]


@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p)[:-3] for p in GOLDEN])
def test_learn_matches_reference(path):
    fx = load(path)
    cfg = O.config_from_reference_kwargs(**fx['model_kwargs'])
    ref = fx['out']
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and k in ref['grads']) for k, v in fx['state_dict'].items()}
    exp = O.OracleExperience(
        latents=ref['latents'], agent_embed=ref['agent_embed'], rewards=ref['rewards'], values=ref['values'],
        actions=ref['actions'], log_probs=ref['log_probs'], lens=ref['lens'], is_truncated=ref['is_truncated'],
        terminals=ref['terminals'], step_size=ref['step_size'])
    pl, vl, _ = O.learn_from_experience(sd, cfg, exp)
    torch.testing.assert_close(pl.detach(), ref['policy_loss'], atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(vl.detach(), ref['value_loss'], atol=1e-6, rtol=1e-5)
    (pl + vl).backward()
    for name, g in ref['grads'].items():
        torch.testing.assert_close(sd[name].grad, g, atol=1e-6, rtol=1e-4, msg=lambda m, n=name: f'{n}: {m}')


@pytest.mark.parametrize('objective', ['ppo', 'spo', 'pmpo'])
@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p)[:-3] for p in GOLDEN])
def test_learn_offpolicy_objectives_match_reference(path, objective):
    """The three surrogate objectives (D4:6127-6212) replayed after the policy head moved: ratio != 1, the PPO clip
    engages and the PMPO KL term against old_action_unembeds is non-zero."""
    fx = load(path)
    cfg = O.config_from_reference_kwargs(**fx['model_kwargs'])
    ref = fx['out']
    want = ref[f'offpolicy_{objective}']
    state = dict(fx['state_dict'])
    state.update(ref['offpolicy_params'])
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and k in want['grads']) for k, v in state.items()}
    exp = O.OracleExperience(
        latents=ref['latents'], agent_embed=ref['agent_embed'], rewards=ref['rewards'], values=ref['values'],
        actions=ref['actions'], log_probs=ref['log_probs'], lens=ref['lens'], is_truncated=ref['is_truncated'],
        terminals=ref['terminals'], step_size=ref['step_size'], old_action_unembeds=ref['old_action_unembeds'])
    pl, vl, _ = O.learn_from_experience(sd, cfg, exp, objective=objective)
    torch.testing.assert_close(pl.detach(), want['policy_loss'], atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(vl.detach(), want['value_loss'], atol=1e-6, rtol=1e-5)
    (pl + vl).backward()
    for name, g in want['grads'].items():
        torch.testing.assert_close(sd[name].grad, g, atol=1e-6, rtol=1e-4, msg=lambda m, n=name: f'{n}: {m}')
    if objective != 'ppo':      # the fixtures must actually separate the objectives
        assert (want['policy_loss'] - ref['offpolicy_ppo']['policy_loss']).abs() > 1e-3


PROMPTED = [p for p in GOLDEN if 'prompted' in load(p)['out']]


@pytest.mark.parametrize('flow', ['resume', 'cold', 'env_step'])
@pytest.mark.parametrize('path', PROMPTED, ids=[os.path.basename(p)[:-3] for p in PROMPTED])
def test_prompted_generate_matches_reference(path, flow):
    """generate(prompt_latents=..., prompt_discrete_actions=..., prompt_rewards=...[, time_cache=...]) (D4:6377-6402):
    resumed over the cache, rebuilt cold from the prompt (the reference's uncached multi-frame forward vs the oracle's
    per-frame prefill), and the env wrapper's single step with a caller-supplied action (env.py:464-484)."""
    fx = load(path)
    cfg = O.config_from_reference_kwargs(**fx['model_kwargs'])
    pr = fx['out']['prompted']
    head, ref = pr['head'], pr[flow]
    B = fx['gen_kwargs']['batch_size']
    torch.manual_seed(ref['seed'])
    exp = O.generate(
        fx['state_dict'], cfg, ref['latents'].shape[1], B,
        prompt_latents=head['latents'], prompt_actions=head['actions'], prompt_rewards=head['rewards'],
        kv_cache=None if flow == 'cold' else O.cache_from_reference(head['kv_cache']),
        return_agent_actions=flow != 'env_step')
    tol = dict(atol=2e-5, rtol=1e-4)
    torch.testing.assert_close(exp.latents, ref['latents'], **tol)
    assert torch.equal(exp.latents[:, :pr['P']], head['latents'])                 # prompt frames pass through
    torch.testing.assert_close(exp.agent_embed, ref['agent_embed'], **tol)
    torch.testing.assert_close(exp.rewards, ref['rewards'], **tol)
    torch.testing.assert_close(exp.episode_return, ref['episode_return'], **tol)
    assert torch.equal(exp.lens, ref['lens'])
    if flow == 'env_step':
        assert ref['actions'] is None and exp.actions is None
    else:
        assert torch.equal(exp.actions, ref['actions'])
        torch.testing.assert_close(exp.log_probs, ref['log_probs'], **tol)
        torch.testing.assert_close(exp.values, ref['values'], **tol)
        torch.testing.assert_close(exp.old_action_unembeds, ref['old_action_unembeds'], **tol)
    kv = torch.stack([torch.stack(layer) for layer in exp.kv_cache])
    assert kv.shape == ref['kv_cache'].shape
    torch.testing.assert_close(kv, ref['kv_cache'], **tol)
    assert ref['token_count'] == exp.latents.shape[1]


def test_cache_continuation_matches_reference():
    """generate(time_cache=...) without prompt latents (the reference's tests/test_dreamer.py::test_cache_generate): each call imagines
    time_steps NEW frames on top of the cached ones.  Golden: three chained calls of the reference (oracle/make_golden_cache.py)."""
    import os
    fx = torch.load(os.path.join(os.path.dirname(__file__), 'golden', 'cache', 'cache_continue.pt'), map_location='cpu', weights_only=False)
    cfg = O.config_from_reference_kwargs(**fx['model_kwargs'])
    cache = None
    for call in fx['calls']:
        torch.manual_seed(call['seed'])
        exp = O.generate(fx['state_dict'], cfg, call['time_steps'], 2, kv_cache=cache)
        cache = exp.kv_cache
        assert torch.equal(exp.actions, call['actions'])
        for name in ('latents', 'rewards', 'values', 'log_probs', 'agent_embed'):
            torch.testing.assert_close(getattr(exp, name), call[name], atol=2e-5, rtol=1e-4, msg=lambda m, n=name: f'{n}: {m}')
        assert cache[0][0].shape[-2] == call['token_count']
        ref_kv = O.cache_from_reference(call['kv'])
        for (k, v), (rk, rv) in zip(cache, ref_kv):
            torch.testing.assert_close(k, rk, atol=2e-5, rtol=1e-4)
            torch.testing.assert_close(v, rv, atol=2e-5, rtol=1e-4)


def test_learn_with_reward_ema_stats_matches_reference():
    """keep_reward_ema_stats=True (D4:5987-6013): quantile-filtered running mean / variance of the returns normalise returns and old
    values before the advantage.  Golden: two consecutive updates of the reference (oracle/make_golden_learn_ema.py)."""
    fx = load(os.path.join(os.path.dirname(__file__), 'golden', 'learn', 'learn_ema.pt'))
    mk = dict(fx['model_kwargs'])
    cfg = O.config_from_reference_kwargs(**mk)
    e = fx['experience']
    exp = O.OracleExperience(latents=e['latents'], agent_embed=e['agent_embed'], rewards=e['rewards'], values=e['values'], actions=e['actions'],
                             log_probs=e['log_probs'], lens=e['lens'], is_truncated=e['is_truncated'], terminals=e['terminals'],
                             step_size=e['step_size'], old_action_unembeds=e['old_action_unembeds'])
    ema = dict(mean=torch.tensor(0.), var=torch.tensor(1.), decay=mk['reward_ema_decay'], quantiles=mk['reward_quantile_filter'])
    for call in fx['calls']:
        sd = {k: v.clone().requires_grad_(v.is_floating_point() and k in call['grads']) for k, v in fx['state_dict'].items()}
        pl, vl, _ = O.learn_from_experience(sd, cfg, exp, objective=call['objective'], ema_stats=ema)
        torch.testing.assert_close(ema['mean'], call['ema_returns_mean'], atol=1e-6, rtol=1e-5)
        torch.testing.assert_close(ema['var'], call['ema_returns_var'], atol=1e-6, rtol=1e-5)
        torch.testing.assert_close(pl.detach(), call['policy_loss'], atol=1e-6, rtol=1e-5)
        torch.testing.assert_close(vl.detach(), call['value_loss'], atol=1e-6, rtol=1e-5)
        (pl + vl).backward()
        for name, g in call['grads'].items():
            torch.testing.assert_close(sd[name].grad, g, atol=1e-6, rtol=1e-4, msg=lambda m, n=name: f'{n}: {m}')
