"""GPU parity of DynamicsWorldModel.interact_with_env (reference dreamer4.py:5470-5889) through d4_observe, against the oracle on
the deterministic toy env (oracle/toy_env.py) with observations tokenized by the oracle's incremental tokenizer through
`obs_to_latents_fn` (there is no CUDA tokenizer yet).

First hardware run: profiles/r1_interact_tests.log (4 cases green on a B200, together with the d4_frame golden tests after the
frame_impl refactor)."""
import os

import pytest
import torch

from oracle import dreamer4_oracle as O
from oracle import tokenizer_oracle as TO
from oracle.toy_env import ToyImageEnv

pytestmark = pytest.mark.gpu

WORLD = os.path.join(os.path.dirname(__file__), 'golden', 'tokenizer', 'world_with_tokenizer.pt')
TOL = dict(atol=5e-5, rtol=2e-4)


@pytest.mark.parametrize('case', range(4), ids=['vec_mixed_bootstrap', 'vec_all_terminated', 'single_truncated', 'single_terminated'])
def test_interact_with_env_matches_oracle(case):
    from dreamer4_b200 import DynamicsWorldModel
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
    model = model.cuda()
    B = 3 if vectorized else 1

    def obs_to_latents(world_model, obs, cache):
        frame = obs['image'].cpu()
        frame = frame if vectorized else frame[None]
        tok_cache, t = cache if cache is not None else (None, 0)
        lat, tok_cache = TO.tokenize_step(tsd, tcfg, frame, tok_cache, t)
        return lat[:, None].cuda(), (tok_cache, t + 1)

    torch.manual_seed(7)
    exp = model.interact_with_env(ToyImageEnv(batch=B if vectorized else None, terminate_at=terminate_at), max_timesteps=max_timesteps,
                                  env_is_vectorized=vectorized, obs_to_latents_fn=obs_to_latents)
    torch.manual_seed(7)                                                       # the sampler's draws, in interact_with_env's order
    draws = torch.stack([torch.cat([torch.rand(B, n, device='cuda') for n in model.cfg.num_discrete_actions], dim=-1)
                         for _ in range(max_timesteps)]).cpu()
    ref = O.interact_with_env(fx['state_dict'], ocfg, (tsd, tcfg), ToyImageEnv(batch=B if vectorized else None, terminate_at=terminate_at),
                              max_timesteps=max_timesteps, env_is_vectorized=vectorized, noise=O.InjectedNoise(None, draws, None))
    assert torch.equal(exp.actions.discrete.cpu(), ref.actions)
    for name in ('lens', 'terminals', 'is_truncated'):
        assert torch.equal(getattr(exp, name).cpu(), getattr(ref, name)), name
    for name in ('latents', 'agent_embed', 'rewards', 'values', 'episode_return'):
        torch.testing.assert_close(getattr(exp, name).cpu(), getattr(ref, name), **TOL, msg=lambda m, n=name: f'{n}: {m}')
    torch.testing.assert_close(exp.log_probs.discrete.cpu(), ref.log_probs, **TOL)
    torch.testing.assert_close(exp.old_action_unembeds.discrete.cpu(), ref.old_action_unembeds, **TOL)
    assert exp.video.shape[2] == exp.rewards.shape[1] and not exp.is_from_world_model
