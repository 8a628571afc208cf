"""generate(time_cache=...) without prompt latents on the GPU (the reference's tests/test_dreamer.py::test_cache_generate flow): three
chained calls against the oracle on the same injected draws; the oracle is pinned to the reference's own chained calls
(tests/golden/cache/cache_continue.pt, tests/test_oracle_golden.py).

Green on hardware since the driver's round-1 run; strict since round 2."""
import os

import pytest
import torch

from oracle import dreamer4_oracle as O

pytestmark = pytest.mark.gpu


def test_cache_continuation_matches_oracle():
    from dreamer4_b200 import DynamicsWorldModel
    import test_gpu_parity as G
    fx = torch.load(os.path.join(os.path.dirname(__file__), 'golden', 'cache', 'cache_continue.pt'), map_location='cpu', weights_only=False)
    model = DynamicsWorldModel(**fx['model_kwargs'], precision='fp32')
    model.load_state_dict(fx['state_dict'], strict=True)
    model = model.cuda()
    ocfg = O.config_from_reference_kwargs(**fx['model_kwargs'])
    tc, ocache = None, None
    for i, call in enumerate(fx['calls']):
        T = call['time_steps']
        noise = G.make_noise(model.cfg, T, 2, seed=40 + i)
        ref = O.generate(fx['state_dict'], ocfg, T, 2, noise=O.InjectedNoise(noise['latent'], noise['action_uniform'], noise['terminal_uniform']), kv_cache=ocache)
        ocache = ref.kv_cache
        exp, tc = model.generate(T, batch_size=2, noise=G.to_cuda(noise), time_cache=tc, return_time_cache=True, return_rewards_per_frame=True,
                                 return_agent_actions=True, return_log_probs_and_values=True)
        assert tc.main.token_count == call['token_count']
        ref_kv = torch.stack([torch.stack(layer) for layer in ref.kv_cache])
        G.compare_experience(exp, ref, tc.main.next_kv_cache, ref_kv)
