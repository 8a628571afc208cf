"""Golden vectors for the inference branch of DynamicsWorldModel.forward (reference dreamer4.py:6792-7295), made by executing the
reference's own source: a 4-frame parallel call with per-dream signal levels / step sizes and discrete actions, the same frames fed
one by one through the time cache (the reference's tests/test_dreamer.py::test_e2e_sequential_parallel_cache flow), and the policy /
value heads called as modules on the agent embeddings.  Build-container only.

    python oracle/make_golden_forward.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_import import import_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden', 'forward', 'forward_inference.pt')
MODEL = dict(dim=32, dim_latent=8, num_latent_tokens=6, depth=4, time_block_every=2, attn_heads=2, attn_dim_head=16,
             num_discrete_actions=(3, 4), predict_terminals=False, num_tasks=3)


def main(seed=41):
    ref = import_reference()
    torch.manual_seed(seed)
    model = ref.DynamicsWorldModel(**MODEL).eval()
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith('gamma') or n.endswith('norm.weight') or n.endswith('norm_context.weight') or '.1.weight' in n:
                p.add_(torch.randn_like(p) * 0.1)
            if 'unembed' in n or n.endswith('queries') or 'learned_embed' in n or n == 'register_tokens':
                p.mul_(30.)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    B, T = 3, 4
    latents = torch.randn(B, T, 6, 8).clamp(-1, 1)
    signal_levels = torch.randint(0, 64, (B, T))
    step_sizes = torch.tensor([1., 4., 16.])
    actions = torch.stack((torch.randint(0, 3, (B, T)), torch.randint(0, 4, (B, T))), dim=-1)
    tasks = torch.tensor([0, 2, 1])
    with torch.no_grad():
        pred, (embeds, inter) = model(latents=latents, signal_levels=signal_levels, step_sizes=step_sizes, discrete_actions=actions, tasks=tasks,
                                      return_pred_only=True, return_intermediates=True, latent_is_noised=True)
        policy_embed = model.policy_head(embeds.agent)
        value_bins = model.value_head(embeds.agent)
        seq_flow, seq_agent, cache = [], [], None
        for i in range(T):
            act = None if i == 0 else actions[:, i - 1:i]
            p_i, (e_i, cache) = model(latents=latents[:, i:i + 1], signal_levels=signal_levels[:, i:i + 1], step_sizes=step_sizes, discrete_actions=act,
                                      tasks=tasks, time_cache=cache, return_pred_only=True, return_intermediates=True, latent_is_noised=True)
            seq_flow.append(p_i.flow)
            seq_agent.append(e_i.agent)
        seq_flow, seq_agent = torch.cat(seq_flow, dim=1), torch.cat(seq_agent, dim=1)
    print('parallel vs sequential: flow', float((pred.flow - seq_flow).abs().max()), 'agent', float((embeds.agent - seq_agent).abs().max()))
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    torch.save(dict(model_kwargs=MODEL, state_dict=sd, latents=latents, signal_levels=signal_levels, step_sizes=step_sizes, actions=actions, tasks=tasks,
                    flow=pred.flow.clone(), agent=embeds.agent.clone(), kv_cache=inter.main.next_kv_cache.clone(), token_count=inter.main.token_count,
                    policy_embed=policy_embed.clone(), value_bins=value_bins.clone(), seq_flow=seq_flow.clone(), seq_agent=seq_agent.clone(),
                    seq_kv_cache=cache.main.next_kv_cache.clone(), torch_version=torch.__version__), OUT)
    print('flow', tuple(pred.flow.shape), 'agent', tuple(embeds.agent.shape), 'kv', tuple(inter.main.next_kv_cache.shape), '->', OUT, os.path.getsize(OUT) / 1e6, 'MB')


if __name__ == '__main__':
    main()
