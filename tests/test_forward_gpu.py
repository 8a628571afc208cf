"""DynamicsWorldModel.forward, inference branch (reference dreamer4.py:6792-7295), and the head modules called directly, on the GPU
against golden vectors of the reference itself (oracle/make_golden_forward.py): a 4-frame parallel call with per-dream signal levels /
step sizes, discrete actions and tasks; the same frames fed one at a time through the returned time cache (the reference's
tests/test_dreamer.py::test_e2e_sequential_parallel_cache flow); `model.policy_head(embeds.agent)` / `model.value_head(...)`."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = dict(atol=5e-5, rtol=2e-4)
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'forward', 'forward_inference.pt')


def build(precision='fp32'):
    from dreamer4_b200 import DynamicsWorldModel
    fx = torch.load(GOLDEN, map_location='cpu', weights_only=False)
    model = DynamicsWorldModel(**fx['model_kwargs'], precision=precision)
    model.load_state_dict(fx['state_dict'], strict=True)
    return fx, model.cuda().eval()


def test_parallel_forward_matches_reference_golden():
    fx, model = build()
    cu = lambda k: fx[k].cuda()
    pred, (embeds, inter) = model(latents=cu('latents'), signal_levels=cu('signal_levels'), step_sizes=cu('step_sizes'), discrete_actions=cu('actions'),
                                  tasks=cu('tasks'), return_pred_only=True, return_intermediates=True, latent_is_noised=True)
    assert pred.flow.shape == fx['flow'].shape and embeds.agent.shape == fx['agent'].shape
    torch.testing.assert_close(pred.flow.cpu(), fx['flow'], **TOL)
    torch.testing.assert_close(embeds.agent.cpu(), fx['agent'], **TOL)
    assert inter.main.token_count == fx['token_count'] and inter.main.next_kv_cache.shape == fx['kv_cache'].shape
    torch.testing.assert_close(inter.main.next_kv_cache.cpu(), fx['kv_cache'], **TOL)
    # the heads as callables (reference: nn.Modules; tests/test_dreamer.py:1262)
    torch.testing.assert_close(model.policy_head(embeds.agent).cpu(), fx['policy_embed'], **TOL)
    torch.testing.assert_close(model.value_head(embeds.agent).cpu(), fx['value_bins'], **TOL)
    # without intermediates only the prediction comes back
    only = model(latents=cu('latents'), signal_levels=cu('signal_levels'), step_sizes=cu('step_sizes'), discrete_actions=cu('actions'), tasks=cu('tasks'),
                 return_pred_only=True, latent_is_noised=True)
    assert torch.equal(only.flow, pred.flow)


def test_sequential_forward_over_the_time_cache_matches_parallel():
    """tests/test_dreamer.py:1206-1296: policy embeds of the parallel call and of frame-by-frame calls with `time_cache` agree (atol 1e-4
    there); here additionally both match the reference's own numbers."""
    fx, model = build()
    cu = lambda k: fx[k].cuda()
    T = fx['latents'].shape[1]
    flows, agents, cache = [], [], None
    for i in range(T):
        act = None if i == 0 else cu('actions')[:, i - 1:i]
        p_i, (e_i, cache) = model(latents=cu('latents')[:, i:i + 1], signal_levels=cu('signal_levels')[:, i:i + 1], step_sizes=cu('step_sizes'),
                                  discrete_actions=act, tasks=cu('tasks'), time_cache=cache, return_pred_only=True, return_intermediates=True,
                                  latent_is_noised=True)
        flows.append(p_i.flow.clone())
        agents.append(e_i.agent.clone())
    flow, agent = torch.cat(flows, dim=1), torch.cat(agents, dim=1)
    torch.testing.assert_close(flow.cpu(), fx['seq_flow'], **TOL)
    torch.testing.assert_close(agent.cpu(), fx['seq_agent'], **TOL)
    torch.testing.assert_close(cache.main.next_kv_cache.cpu(), fx['seq_kv_cache'], **TOL)
    par = model(latents=cu('latents'), signal_levels=cu('signal_levels'), step_sizes=cu('step_sizes'), discrete_actions=cu('actions'), tasks=cu('tasks'),
                return_pred_only=True, return_intermediates=True, latent_is_noised=True)[1][0].agent
    assert torch.allclose(model.policy_head(par), model.policy_head(agent), atol=1e-4)


TRAIN = os.path.join(os.path.dirname(__file__), 'golden', 'forward', 'forward_training.pt')


@pytest.mark.parametrize('precision', ['fp32', 'tf32x3'])
@pytest.mark.parametrize('case', [0, 1, 2, 3], ids=['shortcut', 'shortcut_var_len', 'plain', 'plain_var_len'])
def test_training_forward_losses_match_reference_golden(case, precision):
    """The TRAINING branch of forward(), forward only (reference dreamer4.py:6963-6997, 7297-7743): flow, shortcut, multi-token reward /
    discrete-action and terminal losses and their total against the reference's own numbers on the schedule and noise its seed produced
    (oracle/make_golden_training_forward.py)."""
    from dreamer4_b200 import DynamicsWorldModel
    fx = torch.load(TRAIN, map_location='cpu', weights_only=False)
    ref = fx['cases'][case]
    model = DynamicsWorldModel(**fx['model_kwargs'], precision=precision)
    model.load_state_dict(fx['state_dict'], strict=True)
    model = model.cuda()
    cu = lambda t: t.cuda()
    kw = dict(latents=cu(fx['latents']), rewards=cu(fx['rewards']), discrete_actions=cu(fx['actions']), terminals=cu(fx['terminals']), tasks=cu(fx['tasks']),
              return_all_losses=True, noise=cu(ref['noise']),
              train_schedule=dict(shortcut_train=ref['shortcut_train'], step_sizes_log2=cu(ref['step_sizes_log2']), signal_levels=cu(ref['signal_levels'])))
    if ref['var_len']:
        kw['lens'] = cu(fx['lens'])
    total, losses = model(**kw)
    rtol = 2e-4 if precision == 'fp32' else 1e-3
    for name in ('flow', 'shortcut', 'rewards', 'terminals', 'discrete_actions'):
        torch.testing.assert_close(getattr(losses, name).cpu(), ref['losses'][name], atol=5e-6, rtol=rtol, msg=lambda m, n=name: f'{n}: {m}')
    torch.testing.assert_close(total.cpu(), ref['total'], atol=1e-5, rtol=rtol)


def test_training_forward_draws_its_own_schedule():
    """Without injected draws the branch samples the shortcut coin, step sizes, signal levels and noise itself (seeded): finite, repeatable."""
    fx, model = build()
    kw = dict(latents=fx['latents'].cuda(), discrete_actions=fx['actions'].cuda(), tasks=fx['tasks'].cuda(), seed=7)
    a, b = model(**kw), model(**kw)
    assert a.ndim == 0 and bool(torch.isfinite(a)) and torch.equal(a, b)
