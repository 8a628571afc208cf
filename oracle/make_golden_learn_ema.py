"""Golden vectors for learn_from_experience with keep_reward_ema_stats=True (reference dreamer4.py:5987-6013), made by executing
the reference's own source (third-party deps shimmed, see make_golden.py).  Two consecutive updates on the same dream so that
the running statistics move: per call the losses, head gradients and the EMA buffers.  Build-container only.

    python oracle/make_golden_learn_ema.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_import import import_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden', 'learn', 'learn_ema.pt')
MODEL = dict(dim=32, dim_latent=8, num_latent_tokens=6, depth=2, time_block_every=2, attn_heads=2, attn_dim_head=16,
             num_discrete_actions=(3, 4), predict_terminals=False, keep_reward_ema_stats=True, reward_ema_decay=0.9,
             reward_quantile_filter=(0.1, 0.9))


def main(seed=31):
    ref = import_reference()
    torch.manual_seed(seed)
    model = ref.DynamicsWorldModel(**MODEL)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith('gamma') or n.endswith('norm.weight') or n.endswith('norm_context.weight') or '.1.weight' in n:
                p.add_(torch.randn_like(p) * 0.1)
            if 'unembed' in n or n.endswith('queries') or 'learned_embed' in n or n == 'register_tokens':
                p.mul_(30.)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    torch.manual_seed(seed + 1)
    exp = model.generate(time_steps=7, batch_size=5, return_rewards_per_frame=True, return_agent_actions=True, return_log_probs_and_values=True)
    # a spread of returns: the random-init reward head predicts ~0 everywhere
    exp.rewards = exp.rewards + torch.randn_like(exp.rewards) * 2.
    exp.lens = torch.tensor([7, 5, 7, 3, 6])
    exp.is_truncated = torch.tensor([True, False, True, True, False])
    fields = dict(latents=exp.latents, agent_embed=exp.agent_embed, rewards=exp.rewards, values=exp.values, actions=exp.actions.discrete,
                  log_probs=exp.log_probs.discrete, old_action_unembeds=exp.old_action_unembeds.discrete, lens=exp.lens,
                  is_truncated=exp.is_truncated, terminals=exp.terminals, step_size=exp.step_size)
    calls = []
    for objective in ('ppo', 'spo'):
        model.zero_grad()
        pl, vl = model.learn_from_experience(exp, objective=objective)
        pl.backward(retain_graph=True)
        vl.backward()
        calls.append(dict(objective=objective, policy_loss=pl.detach().clone(), value_loss=vl.detach().clone(),
                          grads={n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None},
                          ema_returns_mean=model.ema_returns_mean.clone(), ema_returns_var=model.ema_returns_var.clone()))
        print(objective, float(pl), float(vl), float(model.ema_returns_mean), float(model.ema_returns_var))
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    torch.save(dict(model_kwargs=MODEL, state_dict=sd, experience={k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in fields.items()},
                    calls=calls, torch_version=torch.__version__), OUT)
    print('->', OUT, os.path.getsize(OUT) / 1e6, 'MB')


if __name__ == '__main__':
    main()
