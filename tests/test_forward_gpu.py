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


def test_training_branch_is_refused_loudly():
    fx, model = build()
    with pytest.raises(NotImplementedError):
        model(latents=fx['latents'].cuda())
