"""Is a config-4 rollout bit-reproducible run to run?  Runs generate() twice per setting on the same injected noise and compares
every output bit for bit; prints one JSON line per setting (first differing frame per field, number of dreams whose sampled
actions differ).  Settings: engine precision x D4_FUSE_SS (the only atomics on the rollout path are the fused sums of squares).

    python scripts/determinism_check.py [--batch 2048] [--horizon 12]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import WORKLOADS  # noqa: E402
from dreamer4_b200 import DynamicsWorldModel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=2048)
    ap.add_argument('--horizon', type=int, default=12)
    ap.add_argument('--precisions', default='tf32x3,f16x3')
    args = ap.parse_args()
    B, H = args.batch, args.horizon
    cfgm = WORKLOADS['config4']['model']
    g = torch.Generator(device='cuda').manual_seed(3)
    noise = dict(latent=torch.randn(H, B, cfgm['num_latent_tokens'], cfgm['dim_latent'], device='cuda', generator=g),
                 action_uniform=torch.rand(H, B, cfgm['num_discrete_actions'], device='cuda', generator=g),
                 terminal_uniform=torch.rand(H, B, device='cuda', generator=g))
    flags = dict(return_rewards_per_frame=True, return_agent_actions=True, return_log_probs_and_values=True)
    for prec in args.precisions.split(','):
        for fuse in ('1', '0'):
            os.environ['D4_FUSE_SS'] = fuse
            torch.manual_seed(0)
            model = DynamicsWorldModel(**cfgm, precision=prec)
            with torch.no_grad():
                for n, p in model.named_parameters():
                    if 'unembed' in n:
                        p.mul_(30.)
            model = model.cuda()
            runs = []
            for _ in range(2):
                e = model.generate(H, batch_size=B, noise=noise, **flags)
                runs.append({k: getattr(e, k).clone() for k in ('latents', 'rewards', 'values', 'agent_embed')} |
                            dict(actions=e.actions.discrete.clone(), log_probs=e.log_probs.discrete.clone(), logits=e.old_action_unembeds.discrete.clone()))
            a, b = runs
            first = {}
            for k in a:
                neq = (a[k] != b[k]).flatten(2).any(dim=2) if a[k].ndim > 2 else (a[k] != b[k])       # (B, T)
                per_t = neq.any(dim=0)
                first[k] = int(per_t.nonzero()[0]) if bool(per_t.any()) else None
            forked = int((a['actions'] != b['actions']).flatten(1).any(dim=1).sum())
            print(json.dumps(dict(check='rerun bit-identical', precision=prec, fuse_ss=fuse, batch=B, horizon=H,
                                  identical=all(v is None for v in first.values()), first_differing_frame=first, forked_dreams=forked)), flush=True)
            del model
            torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
