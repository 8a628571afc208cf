"""Generate golden vectors by executing the reference's OWN source (test infrastructure).

Runs /root/reference/dreamer4/dreamer4.py (unmodified; third-party deps replaced by oracle/shims)
on small seeded cases and stores inputs + outputs under tests/golden/*.pt.  Build-container only:
the GPU box has no /root/reference, so the fixtures are committed.

    python oracle/make_golden.py            # regenerates every fixture
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_import import import_reference  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), 'tests', 'golden')

# name -> (DynamicsWorldModel kwargs, generate kwargs)
CASES = {
    # default architecture shrunk: LQAP pools both sides, 1 time layer of 4, single discrete action
    'tiny_default': (
        dict(dim=32, dim_latent=8, num_latent_tokens=6, depth=4, time_block_every=4, attn_heads=2, attn_dim_head=16,
             num_discrete_actions=4, predict_terminals=False),
        dict(time_steps=6, batch_size=3),
    ),
    # two time layers, multi-discrete actions, GQA, same-length latent/spatial tokens (Linear in/out)
    'tiny_gqa_multidiscrete': (
        dict(dim=32, dim_latent=16, num_latent_tokens=2, num_spatial_tokens=2, depth=4, time_block_every=2,
             attn_heads=2, attn_dim_head=16, attn_kwargs=dict(query_heads=4), num_discrete_actions=(3, 5),
             num_register_tokens=3, predict_terminals=False,
             reward_encoder_kwargs=dict(num_bins=20, reward_range=(-3., 3.)), value_encoder_kwargs=dict(num_bins=32)),
        dict(time_steps=5, batch_size=2),
    ),
    # every layer a time layer + terminal head with early lens bookkeeping + GEGLU feed-forward
    'tiny_terminals_geglu': (
        dict(dim=32, dim_latent=8, num_latent_tokens=4, depth=2, time_block_every=1, attn_heads=2, attn_dim_head=16,
             num_discrete_actions=4, predict_terminals=True, ff_kwargs=dict(activation='gelu')),
        dict(time_steps=5, batch_size=4, return_terminals=True),
    ),
}


# name -> VideoTokenizer kwargs (lpips off: its network is a torchvision download)
TOKENIZER_CASES = {
    # square image, one time layer of two per stack, the default single flow step of the decoder
    'tokenizer_tiny': dict(dim=32, dim_latent=8, patch_size=4, image_size=16, num_latent_tokens=6, encoder_depth=2, decoder_depth=2,
                           time_block_every=2, attn_heads=2, attn_dim_head=16, lpips_loss_weight=0.),
    # non-square image, every layer a time layer, GEGLU, two decoder flow steps, one channel
    'tokenizer_rect_flow2': dict(dim=32, dim_latent=4, patch_size=4, image_height=8, image_width=16, num_latent_tokens=3, encoder_depth=2,
                                 decoder_depth=3, time_block_every=1, attn_heads=2, attn_dim_head=16, channels=1, decoder_flow_steps=2,
                                 ff_kwargs=dict(activation='gelu'), lpips_loss_weight=0.),
}


def run_tokenizer_case(ref, name, kwargs, seed=17, batch=2, frames=3):
    """VideoTokenizer.tokenize and .decode (D4:4107-4113, 4183-4237) on a seeded random video."""
    torch.manual_seed(seed)
    tok = ref.VideoTokenizer(**kwargs)
    with torch.no_grad():
        for n, p in tok.named_parameters():
            if n.endswith('gamma') or ('norm' in n and n.endswith('weight')):
                p.add_(torch.randn_like(p) * 0.1)
            if n == 'latent_tokens':
                p.mul_(30.)
    tok.eval()
    sd = {k: v.detach().clone() for k, v in tok.state_dict().items()}
    channels = kwargs.get('channels', 3)
    video = torch.rand(batch, channels, frames, tok.image_height, tok.image_width)
    latents = tok.tokenize(video)
    decode_seed = seed + 1
    torch.manual_seed(decode_seed)
    recon = tok.decode(latents)
    image_latents = tok.tokenize(video[:, :, 0])          # (b c h w) input
    fixture = dict(name=name, tokenizer_kwargs=kwargs, state_dict=sd, video=video, latents=latents.detach().clone(),
                   decode_seed=decode_seed, recon=recon.detach().clone(), image_latents=image_latents.detach().clone(),
                   torch_version=torch.__version__)
    os.makedirs(os.path.join(GOLDEN_DIR, 'tokenizer'), exist_ok=True)     # a directory of its own: tests glob golden/*.pt for the dynamics cases
    path = os.path.join(GOLDEN_DIR, 'tokenizer', f'{name}.pt')
    torch.save(fixture, path)
    print(f'{name}: latents {tuple(latents.shape)} recon {tuple(recon.shape)} -> {path} ({os.path.getsize(path) / 1e6:.2f} MB)')


def run_world_with_tokenizer(ref, name='world_with_tokenizer', seed=23):
    """DynamicsWorldModel with its VideoTokenizer attached: generate(prompt=video) -> video (D4:6377-6387, 6699-6724) and the
    DreamTrainer-flag rollout with return_decoded_video (Experience.video)."""
    tk = TOKENIZER_CASES['tokenizer_tiny']
    mk = dict(dim=32, dim_latent=8, depth=4, time_block_every=4, attn_heads=2, attn_dim_head=16, num_discrete_actions=4, predict_terminals=False)
    torch.manual_seed(seed)
    tok = ref.VideoTokenizer(**tk)
    model = ref.DynamicsWorldModel(video_tokenizer=tok, **mk)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith('gamma') or ('norm' in n and n.endswith('weight')) or '.1.weight' in n:
                p.add_(torch.randn_like(p) * 0.1)
            if 'unembed' in n or n.endswith('queries') or 'learned_embed' in n or n in ('register_tokens', 'video_tokenizer.latent_tokens'):
                p.mul_(30.)
    model.eval()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    B, T, P = 2, 4, 2
    prompt = torch.rand(B, 3, P, tok.image_height, tok.image_width)
    torch.manual_seed(seed + 1)
    prompted_video = model.generate(time_steps=T, batch_size=B, prompt=prompt)
    torch.manual_seed(seed + 2)
    exp = model.generate(time_steps=3, batch_size=B, return_rewards_per_frame=True, return_agent_actions=True, return_log_probs_and_values=True)
    # interact_with_env (D4:5470-5889) on oracle/toy_env.py: vectorized with one episode terminating early, one never (cut by
    # max_timesteps -> bootstrap branch) and one late; all terminated (no bootstrap); a single non-vectorized env both ways
    from toy_env import ToyImageEnv
    interact = []
    for i, (vectorized, terminate_at) in enumerate(((True, [2, 0, 4]), (True, [1, 2, 3]), (False, None), (False, [3]))):
        env = ToyImageEnv(batch=3 if vectorized else None, terminate_at=terminate_at)
        torch.manual_seed(seed + 10 + i)
        e = model.interact_with_env(env, max_timesteps=5, env_is_vectorized=vectorized)
        interact.append(dict(
            vectorized=vectorized, terminate_at=terminate_at, max_timesteps=5, seed=seed + 10 + i,
            latents=e.latents, agent_embed=e.agent_embed, rewards=e.rewards, values=e.values, actions=e.actions.discrete,
            log_probs=e.log_probs.discrete, old_action_unembeds=e.old_action_unembeds.discrete, lens=e.lens,
            is_truncated=e.is_truncated, terminals=e.terminals, episode_return=e.episode_return, video=e.video))
    interact = [{k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in d.items()} for d in interact]

    fixture = dict(name=name, tokenizer_kwargs=tk, model_kwargs=mk, state_dict=sd, prompt=prompt, interact=interact,
                   prompted=dict(seed=seed + 1, time_steps=T, video=prompted_video.detach().clone()),
                   dream=dict(seed=seed + 2, time_steps=3, video=exp.video.detach().clone(), latents=exp.latents.detach().clone(),
                              actions=exp.actions.discrete.clone(), rewards=exp.rewards.detach().clone()),
                   torch_version=torch.__version__)
    path = os.path.join(GOLDEN_DIR, 'tokenizer', f'{name}.pt')
    torch.save(fixture, path)
    print(f'{name}: prompted video {tuple(prompted_video.shape)} dream video {tuple(exp.video.shape)} -> {path} ({os.path.getsize(path) / 1e6:.2f} MB)')


def _experience_dict(exp, time_cache):
    get = lambda a: None if a is None else a.discrete
    return dict(latents=exp.latents, agent_embed=exp.agent_embed, rewards=exp.rewards, values=exp.values,
                actions=get(exp.actions), log_probs=get(exp.log_probs), old_action_unembeds=get(exp.old_action_unembeds),
                lens=exp.lens, is_truncated=exp.is_truncated, terminals=exp.terminals, episode_return=exp.episode_return,
                kv_cache=time_cache.main.next_kv_cache, token_count=time_cache.main.token_count)


def run_prompted(model, gen_kwargs, seed, P=2):
    """Three continuations of a P-frame rollout, each from its own seed:
    resume   - prompt + the time cache of those frames, the policy keeps acting (DreamTrainer flags)
    cold     - the same prompt without a cache: the reference's uncached multi-frame forward rebuilds the context
    env_step - env.py's DynamicsWorldModelWrapper.step: one new frame from prompt + cache, the action supplied by the caller"""
    B, T = gen_kwargs['batch_size'], gen_kwargs['time_steps']
    flags = dict(return_rewards_per_frame=True, return_agent_actions=True, return_log_probs_and_values=True, return_time_cache=True)
    torch.manual_seed(seed + 3000)
    head, head_cache = model.generate(time_steps=P, batch_size=B, **flags)
    prompt = dict(prompt_latents=head.latents, prompt_discrete_actions=head.actions.discrete, prompt_rewards=head.rewards)
    res = dict(P=P, head_seed=seed + 3000, head=_experience_dict(head, head_cache))
    torch.manual_seed(seed + 3001)
    res['resume'] = dict(seed=seed + 3001, **_experience_dict(*model.generate(time_steps=T, batch_size=B, time_cache=head_cache, **prompt, **flags)))
    torch.manual_seed(seed + 3002)
    res['cold'] = dict(seed=seed + 3002, **_experience_dict(*model.generate(time_steps=T, batch_size=B, **prompt, **flags)))
    torch.manual_seed(seed + 3003)
    res['env_step'] = dict(seed=seed + 3003, **_experience_dict(*model.generate(
        time_steps=P + 1, batch_size=B, time_cache=head_cache, return_rewards_per_frame=True, return_time_cache=True, **prompt)))
    return res


def run_case(ref, name, model_kwargs, gen_kwargs, seed=7):
    torch.manual_seed(seed)
    model = ref.DynamicsWorldModel(**model_kwargs)
    # default init leaves several gains at exactly 0/1 - perturb so every parameter matters
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith('gamma') or n.endswith('norm.weight') or n.endswith('norm_context.weight') or '.1.weight' in n:
                p.add_(torch.randn_like(p) * 0.1)
            if 'unembed' in n or n.endswith('queries') or 'learned_embed' in n or n == 'register_tokens':
                p.mul_(30.)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}

    gen_seed = seed + 1000
    torch.manual_seed(gen_seed)
    exp, time_cache = model.generate(
        return_rewards_per_frame=True, return_agent_actions=True, return_log_probs_and_values=True,
        return_time_cache=True, **gen_kwargs)

    out = dict(
        latents=exp.latents, agent_embed=exp.agent_embed, rewards=exp.rewards, values=exp.values,
        actions=exp.actions.discrete, log_probs=exp.log_probs.discrete,
        old_action_unembeds=exp.old_action_unembeds.discrete, lens=exp.lens, is_truncated=exp.is_truncated,
        terminals=exp.terminals, episode_return=exp.episode_return, step_size=exp.step_size,
        kv_cache=time_cache.main.next_kv_cache, token_count=time_cache.main.token_count,
    )

    # DreamTrainer marks the last frame as bootstrap-only via is_truncated (TR:1422-1423); generate already did.
    model.zero_grad()
    policy_loss, value_loss = model.learn_from_experience(exp)
    policy_loss.backward(retain_graph=True)
    value_loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    out.update(policy_loss=policy_loss.detach(), value_loss=value_loss.detach(), grads=grads)

    # parameter groups the trainers hand to their optimizers (D4:5335-5363)
    names = {id(p): n for n, p in model.named_parameters()}
    out['param_groups'] = {fn: [names[id(p)] for p in getattr(model, fn)()] for fn in ('muon_parameters', 'policy_head_parameters', 'value_head_parameters')}

    # off-policy replay: nudge the policy head (as an optimizer step would) so the importance ratio leaves 1, the PPO
    # clip engages and the PMPO KL term is non-zero, then run all three surrogate objectives (D4:6127-6212)
    torch.manual_seed(seed + 2000)
    moved = {}
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.startswith('policy_head.') or n == 'action_embedder.discrete_action_unembed':
                p.add_(torch.randn_like(p) * p.abs().mean() * 0.15)
                moved[n] = p.detach().clone()
    out['offpolicy_params'] = moved
    for objective in ('ppo', 'spo', 'pmpo'):
        model.zero_grad()
        pl, vl = model.learn_from_experience(exp, objective=objective)
        pl.backward(retain_graph=True)
        vl.backward()
        # the value loss does not depend on the objective: its gradients are stored once, with 'ppo'
        out[f'offpolicy_{objective}'] = dict(
            policy_loss=pl.detach().clone(), value_loss=vl.detach().clone(),
            grads={n: p.grad.detach().clone() for n, p in model.named_parameters()
                   if p.grad is not None and (objective == 'ppo' or n in moved)})

    # prompted / resumed rollouts (D4:6377-6402, env.py:464-484) from the first `P` frames of the rollout above
    model.load_state_dict(sd)             # undo the off-policy nudge
    if not gen_kwargs.get('return_terminals'):
        out['prompted'] = run_prompted(model, gen_kwargs, seed)

    fixture = dict(name=name, model_kwargs=model_kwargs, gen_kwargs=gen_kwargs, gen_seed=gen_seed,
                   state_dict=sd, out={k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in out.items()},
                   torch_version=torch.__version__)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, f'{name}.pt')
    torch.save(fixture, path)
    print(f'{name}: T={exp.latents.shape[1]} lens={exp.lens.tolist()} policy_loss={policy_loss.item():.6f} '
          f'value_loss={value_loss.item():.6f} -> {path} ({os.path.getsize(path) / 1e6:.2f} MB)')


if __name__ == '__main__':
    ref = import_reference()
    for name, (mk, gk) in CASES.items():
        run_case(ref, name, mk, gk)
    for name, kw in TOKENIZER_CASES.items():
        run_tokenizer_case(ref, name, kw)
    run_world_with_tokenizer(ref)
