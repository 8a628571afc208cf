import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import test_gpu_parity as G
from dreamer4_b200 import DynamicsWorldModel
kwargs = G.BASELINE_MODELS['config4_256px']
for precision in ('fp32', 'tf32x3', 'f16x3'):
    for (Tt, Bt) in ((1, 40), (2, 40), (1, 8)):
        runs = []
        for trim in ('1', '0'):
            os.environ['D4_TRIM_FINAL'] = trim
            torch.manual_seed(21)
            model = DynamicsWorldModel(**kwargs, precision=precision).cuda()
            noise = G.to_cuda(G.make_noise(model.cfg, Tt, Bt, seed=5))
            e, tc = model.generate(Tt, batch_size=Bt, return_rewards_per_frame=True, return_agent_actions=True, return_log_probs_and_values=True,
                                   return_time_cache=True, noise=noise)
            runs.append(dict(latents=e.latents.clone(), agent=e.agent_embed.clone(), kv=tc.main.next_kv_cache.clone(), values=e.values.clone()))
            model._release()
        a, b = runs
        print(precision, 'T', Tt, 'B', Bt, {k: (bool(torch.equal(a[k], b[k])), float((a[k] - b[k]).abs().max())) for k in a}, flush=True)
