"""Golden vectors for the TRAINING branch of DynamicsWorldModel.forward (reference dreamer4.py:6963-6997, 7297-7743), forward only:
the individual losses (flow, shortcut, rewards per prediction step, terminals, discrete actions per prediction step) and their total,
made by executing the reference's own source with a seed, together with the random draws that seed produces (the reference re-seeds
before EVERY draw, dreamer4.py:430-460: shortcut coin, step sizes, signal levels, noise) so that a CUDA implementation can be fed the
same schedule and noise.  Two cases: shortcut training on (seed chosen so the coin lands there) and off.  Build-container only.

    python oracle/make_golden_training_forward.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_import import import_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden', 'forward', 'forward_training.pt')
MODEL = dict(dim=32, dim_latent=8, num_latent_tokens=6, depth=4, time_block_every=2, attn_heads=2, attn_dim_head=16,
             num_discrete_actions=(3, 4), predict_terminals=True, num_tasks=3, multi_token_pred_len=3)


def draws(model, seed, B, T, latents):
    """The four draws of the training branch, each made right after torch.manual_seed(seed) as `with_seed` does."""
    torch.manual_seed(seed)
    shortcut = torch.rand(1).item() < model.prob_shortcut_train
    if shortcut:
        torch.manual_seed(seed)
        log2 = torch.randint(1, model.num_step_sizes_log2, (B,))
        torch.manual_seed(seed)
        sig = torch.randint(0, model.max_steps, (B, T)) // (2 ** log2)[:, None] * (2 ** log2)[:, None]
    else:
        log2 = torch.zeros(B, dtype=torch.long)
        torch.manual_seed(seed)
        sig = torch.randint(0, model.max_steps, (B, T))
    torch.manual_seed(seed)
    noise = torch.randn_like(latents)
    return shortcut, log2, sig, noise


def main(seed=53):
    ref = import_reference()
    torch.manual_seed(seed)
    model = ref.DynamicsWorldModel(**MODEL).eval()
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith('gamma') or n.endswith('norm.weight') or n.endswith('norm_context.weight') or '.1.weight' in n or '.0.weight' in n and 'to_reward_pred' in n:
                p.add_(torch.randn_like(p) * 0.1)
            if 'unembed' in n or n.endswith('queries') or 'learned_embed' in n or n == 'register_tokens':
                p.mul_(30.)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    B, T = 3, 5
    latents = torch.randn(B, T, 6, 8).clamp(-1, 1)
    rewards = torch.randn(B, T) * 3.
    actions = torch.stack((torch.randint(0, 3, (B, T)), torch.randint(0, 4, (B, T))), dim=-1)
    terminals = torch.zeros(B, T, dtype=torch.bool)          # per-frame flags (b, t): a 1-D (b,) tensor trips the reference's own conform step (:6904)
    terminals[0, 4] = terminals[2, 2] = True
    tasks = torch.tensor([0, 2, 1])
    lens = torch.tensor([5, 4, 3])
    cases = []
    want = [True, False]
    s = 100
    while want:
        shortcut, log2, sig, noise = draws(model, s, B, T, latents)
        if shortcut in want:
            want.remove(shortcut)
            for var_len in (False, True):
                kw = dict(latents=latents, rewards=rewards, discrete_actions=actions, terminals=terminals, tasks=tasks, seed=s, return_all_losses=True)
                if var_len:
                    kw['lens'] = lens
                with torch.no_grad():
                    total, losses = model(**kw)
                cases.append(dict(seed=s, var_len=var_len, shortcut_train=shortcut, step_sizes_log2=log2, signal_levels=sig, noise=noise, total=total.clone(),
                                  losses={k: (v.clone() if torch.is_tensor(v) else torch.tensor(float(v))) for k, v in losses._asdict().items()}))
                print(s, 'shortcut' if shortcut else 'plain', 'var_len' if var_len else 'full', float(total),
                      {k: [round(x, 5) for x in v.flatten().tolist()] for k, v in cases[-1]['losses'].items() if float(v.abs().sum()) > 0})
        s += 1
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    torch.save(dict(model_kwargs=MODEL, state_dict=sd, latents=latents, rewards=rewards, actions=actions, terminals=terminals, tasks=tasks, lens=lens,
                    cases=cases, torch_version=torch.__version__), OUT)
    print('->', OUT, os.path.getsize(OUT) / 1e6, 'MB')


if __name__ == '__main__':
    main()
