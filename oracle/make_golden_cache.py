"""Golden vectors for generate(time_cache=...) WITHOUT prompt latents - the reference's tests/test_dreamer.py::test_cache_generate
flow (D4:6307-6774 with `time_cache` only): every call imagines `time_steps` NEW frames on top of the frames already in the
cache.  Executes the reference's own source (oracle/ref_import.py), build-container only; the fixture is committed.

    python oracle/make_golden_cache.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_import import import_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden', 'cache', 'cache_continue.pt')
KW = dict(dim=32, dim_latent=8, num_latent_tokens=4, depth=2, time_block_every=1, attn_heads=2, attn_dim_head=16, num_discrete_actions=4,
          predict_terminals=False)
FLAGS = dict(return_rewards_per_frame=True, return_agent_actions=True, return_log_probs_and_values=True, return_time_cache=True)


def main():
    ref = import_reference()
    torch.manual_seed(31)
    model = ref.DynamicsWorldModel(**KW)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith('gamma') or n.endswith('norm.weight') or n.endswith('norm_context.weight'):
                p.add_(torch.randn_like(p) * 0.1)
            if 'unembed' in n or n.endswith('queries') or 'learned_embed' in n or n == 'register_tokens':
                p.mul_(30.)
    model.eval()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    calls, cache = [], None
    for i, steps in enumerate((2, 1, 3)):
        seed = 100 + i
        torch.manual_seed(seed)
        exp, cache = model.generate(steps, batch_size=2, time_cache=cache, **FLAGS)
        calls.append(dict(seed=seed, time_steps=steps, latents=exp.latents, rewards=exp.rewards, values=exp.values, actions=exp.actions.discrete,
                          log_probs=exp.log_probs.discrete, agent_embed=exp.agent_embed, token_count=cache.main.token_count,
                          kv=cache.main.next_kv_cache.clone()))
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    torch.save(dict(model_kwargs=KW, state_dict=sd, calls=calls, torch_version=torch.__version__), OUT)
    print('wrote', OUT, [tuple(c['latents'].shape) for c in calls], [c['token_count'] for c in calls])


if __name__ == '__main__':
    main()
