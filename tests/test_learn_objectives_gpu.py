"""GPU parity of learn_from_experience's three surrogate objectives ('ppo' | 'spo' | 'pmpo', reference dreamer4.py:6127-6212)
on OFF-policy replays: the golden fixtures hold the reference's own losses and gradients after its policy head was moved
(oracle/make_golden.py), so the importance ratio leaves 1, the PPO clip engages, SPO's quadratic term and PMPO's KL to the
stored unembeds are non-zero.  Exact-fp32 engine mode; same tolerances as tests/test_gpu_parity.py::test_learn_matches_reference_golden."""
import glob
import os

import pytest
import torch

from oracle import dreamer4_oracle as O

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden', '*.pt')))
IDS = [os.path.basename(p)[:-3] for p in GOLDEN]
HEADS = ('policy_head.', 'value_head.')
UNEMBED = 'action_embedder.discrete_action_unembed'


def load(path):
    return torch.load(path, map_location='cpu', weights_only=False)


def build_model(fx, **extra):
    from dreamer4_b200 import DynamicsWorldModel
    extra.setdefault('precision', 'fp32')
    model = DynamicsWorldModel(**fx['model_kwargs'], **extra)
    state = dict(fx['state_dict'])
    state.update(fx['out']['offpolicy_params'])
    model.load_state_dict(state, strict=True)
    return model.cuda(), state


def experience(ref):
    from dreamer4_b200 import Actions, Experience
    cu = lambda t: t.cuda()
    return Experience(latents=cu(ref['latents']), agent_embed=cu(ref['agent_embed']), rewards=cu(ref['rewards']), values=cu(ref['values']),
                      actions=Actions(cu(ref['actions']), None), log_probs=Actions(cu(ref['log_probs']), None), lens=cu(ref['lens']),
                      is_truncated=cu(ref['is_truncated']), terminals=cu(ref['terminals']), step_size=ref['step_size'],
                      old_action_unembeds=Actions(cu(ref['old_action_unembeds']), None))


def oracle_experience(ref):
    return O.OracleExperience(
        latents=ref['latents'], agent_embed=ref['agent_embed'], rewards=ref['rewards'], values=ref['values'],
        actions=ref['actions'], log_probs=ref['log_probs'], lens=ref['lens'], is_truncated=ref['is_truncated'],
        terminals=ref['terminals'], step_size=ref['step_size'], old_action_unembeds=ref['old_action_unembeds'])


@pytest.mark.parametrize('objective', ['ppo', 'spo', 'pmpo'])
@pytest.mark.parametrize('path', GOLDEN, ids=IDS)
def test_offpolicy_objective_matches_reference_golden(path, objective):
    fx = load(path)
    model, _ = build_model(fx)
    ref = fx['out']
    want = ref[f'offpolicy_{objective}']
    pl, vl = model.learn_from_experience(experience(ref), objective=objective)
    torch.testing.assert_close(pl.detach().cpu(), want['policy_loss'], atol=1e-6, rtol=1e-4)
    torch.testing.assert_close(vl.detach().cpu(), want['value_loss'], atol=1e-6, rtol=1e-4)
    pl.backward()
    vl.backward()
    params = dict(model.named_parameters())
    for name, g in want['grads'].items():
        assert params[name].grad is not None, name
        torch.testing.assert_close(params[name].grad.cpu(), g, atol=2e-6, rtol=2e-4, msg=lambda m, n=name: f'{n}: {m}')


@pytest.mark.parametrize('kw', [dict(pmpo_reverse_kl=False), dict(pmpo_kl_div_loss_weight=0.), dict(pmpo_pos_to_neg_weight=0.8),
                                dict(use_delight_gating=False)],
                         ids=['forward_kl', 'no_kl', 'alpha', 'no_gate'])
def test_pmpo_variants_match_oracle(kw):
    """The PMPO switches the golden run leaves at their defaults (D4:4734-4736), against the oracle's autograd."""
    fx = load([p for p in GOLDEN if 'multidiscrete' in p][0])
    model, state = build_model(fx, **kw)
    ref = fx['out']
    cfg = O.config_from_reference_kwargs(**fx['model_kwargs'], **kw)
    keys = [k for k in state if k.startswith(HEADS) or k == UNEMBED]
    sd = {k: (v.clone().requires_grad_(True) if k in keys else v) for k, v in state.items()}
    rpl, rvl, _ = O.learn_from_experience(sd, cfg, oracle_experience(ref), objective='pmpo')
    (rpl + rvl).backward()
    pl, vl = model.learn_from_experience(experience(ref), objective='pmpo')
    torch.testing.assert_close(pl.detach().cpu(), rpl.detach(), atol=1e-6, rtol=1e-4)
    pl.backward()
    vl.backward()
    params = dict(model.named_parameters())
    for k in keys:
        torch.testing.assert_close(params[k].grad.cpu(), sd[k].grad, atol=2e-6, rtol=2e-4, msg=lambda m, n=k: f'{n}: {m}')


@pytest.mark.parametrize('objective', ['spo', 'pmpo'])
def test_normalize_advantages_override(objective):
    """normalize_advantages defaults to `objective != 'pmpo'` and can be overridden per call (D4:6021)."""
    fx = load(GOLDEN[0])
    model, state = build_model(fx)
    ref = fx['out']
    cfg = O.config_from_reference_kwargs(**fx['model_kwargs'])
    flip = objective == 'pmpo'          # pmpo with z-scored advantages, spo with raw ones
    rpl, _, aux = O.learn_from_experience(dict(state), cfg, oracle_experience(ref), objective=objective, normalize_advantages=flip)
    pl, _ = model.learn_from_experience(experience(ref), objective=objective, normalize_advantages=flip)
    torch.testing.assert_close(pl.detach().cpu(), rpl.detach(), atol=1e-6, rtol=1e-4)
    torch.testing.assert_close(model.last_learn_aux['advantages'].cpu(), aux['advantage'], atol=1e-5, rtol=1e-4)


def test_unknown_objective_and_missing_unembeds_fail_loudly():
    fx = load(GOLDEN[0])
    model, _ = build_model(fx)
    exp = experience(fx['out'])
    with pytest.raises(ValueError):
        model.learn_from_experience(exp, objective='trpo')
    exp.old_action_unembeds = None
    with pytest.raises(AssertionError):
        model.learn_from_experience(exp, objective='pmpo')


def test_reward_ema_stats_match_reference_golden():
    """keep_reward_ema_stats=True (reference dreamer4.py:5987-6013) through the native path: two consecutive updates of the reference
    on the same dream (oracle/make_golden_learn_ema.py) - losses, head gradients and the running statistics after each."""
    from dreamer4_b200 import Actions, DynamicsWorldModel, Experience
    fx = load(os.path.join(os.path.dirname(__file__), 'golden', 'learn', 'learn_ema.pt'))
    model = DynamicsWorldModel(**fx['model_kwargs'], precision='fp32')
    model.load_state_dict(fx['state_dict'], strict=True)
    model = model.cuda()
    e = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in fx['experience'].items()}
    exp = Experience(latents=e['latents'], agent_embed=e['agent_embed'], rewards=e['rewards'], values=e['values'], actions=Actions(e['actions'], None),
                     log_probs=Actions(e['log_probs'], None), lens=e['lens'], is_truncated=e['is_truncated'], terminals=e['terminals'],
                     step_size=e['step_size'], old_action_unembeds=Actions(e['old_action_unembeds'], None))
    params = dict(model.named_parameters())
    for call in fx['calls']:
        model.zero_grad()
        pl, vl = model.learn_from_experience(exp, objective=call['objective'])
        torch.testing.assert_close(model.ema_returns_mean.cpu(), call['ema_returns_mean'], atol=1e-6, rtol=1e-5)
        torch.testing.assert_close(model.ema_returns_var.cpu(), call['ema_returns_var'], atol=1e-6, rtol=1e-5)
        torch.testing.assert_close(pl.detach().cpu(), call['policy_loss'], atol=1e-6, rtol=1e-4)
        torch.testing.assert_close(vl.detach().cpu(), call['value_loss'], atol=1e-6, rtol=1e-4)
        pl.backward()
        vl.backward()
        for name, g in call['grads'].items():
            torch.testing.assert_close(params[name].grad.cpu(), g, atol=2e-6, rtol=2e-4, msg=lambda m, n=name: f'{n}: {m}')
