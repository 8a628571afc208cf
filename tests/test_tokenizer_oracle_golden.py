"""oracle/tokenizer_oracle.py (VideoTokenizer.tokenize / .decode, the steps either side of the rollout - SURVEY.md section 8f
rank 1) against golden vectors produced by the reference's own source (oracle/make_golden.py).  CPU only.  There is no CUDA
path for the tokenizer yet: this pins the checker first, in the order the hot path itself was built."""
import glob
import os

import pytest
import torch

from oracle import dreamer4_oracle as O
from oracle import tokenizer_oracle as TO

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden', 'tokenizer', 'tokenizer_*.pt')))
WORLD = os.path.join(os.path.dirname(__file__), 'golden', 'tokenizer', 'world_with_tokenizer.pt')
IDS = [os.path.basename(p)[:-3] for p in GOLDEN]
TOL = dict(atol=2e-5, rtol=1e-4)


def load(path):
    return torch.load(path, map_location='cpu', weights_only=False)


def test_fixtures_present():
    assert len(GOLDEN) >= 2


@pytest.mark.parametrize('path', GOLDEN, ids=IDS)
def test_tokenize_matches_reference(path):
    fx = load(path)
    cfg = TO.config_from_reference_kwargs(**fx['tokenizer_kwargs'])
    latents = TO.tokenize(fx['state_dict'], cfg, fx['video'])
    assert latents.shape == fx['latents'].shape
    torch.testing.assert_close(latents, fx['latents'], **TOL)
    assert latents.abs().max() <= 1.                                                       # tanh bottleneck (D4:4426)
    torch.testing.assert_close(TO.tokenize(fx['state_dict'], cfg, fx['video'][:, :, 0]), fx['image_latents'], **TOL)   # (b c h w) input


@pytest.mark.parametrize('path', GOLDEN, ids=IDS)
def test_decode_matches_reference(path):
    fx = load(path)
    cfg = TO.config_from_reference_kwargs(**fx['tokenizer_kwargs'])
    torch.manual_seed(fx['decode_seed'])                                                   # replays the randn at D4:4204
    recon = TO.decode(fx['state_dict'], cfg, fx['latents'])
    assert recon.shape == fx['recon'].shape == fx['video'].shape
    torch.testing.assert_close(recon, fx['recon'], **TOL)


@pytest.mark.parametrize('path', GOLDEN, ids=IDS)
def test_frames_are_causal(path):
    """What lets a decode kernel path run one frame per pass over a time-KV cache: frame t of tokenize / decode depends on
    frames <= t only."""
    fx = load(path)
    cfg = TO.config_from_reference_kwargs(**fx['tokenizer_kwargs'])
    video = fx['video']
    full = TO.tokenize(fx['state_dict'], cfg, video)
    head = TO.tokenize(fx['state_dict'], cfg, video[:, :, :2])
    assert torch.equal(full[:, :2], head)
    noise = torch.randn(video.shape, generator=torch.Generator().manual_seed(3))
    full = TO.decode(fx['state_dict'], cfg, fx['latents'], noise=noise)
    head = TO.decode(fx['state_dict'], cfg, fx['latents'][:, :2], noise=noise[:, :, :2])
    assert torch.equal(full[:, :, :2], head)


def test_patchify_round_trip():
    x = torch.randn(2, 3, 8, 12)
    assert torch.equal(TO.unpatchify(TO.patchify(x, 4), 4, 3, 8, 12), x)


def _world():
    fx = load(WORLD)
    sd = fx['state_dict']
    tsd = {k[len('video_tokenizer.'):]: v for k, v in sd.items() if k.startswith('video_tokenizer.')}
    cfg = O.config_from_reference_kwargs(num_latent_tokens=fx['tokenizer_kwargs']['num_latent_tokens'], **fx['model_kwargs'])
    return fx, sd, cfg, (tsd, TO.config_from_reference_kwargs(**fx['tokenizer_kwargs']))


def test_generate_from_video_prompt_matches_reference():
    """generate(prompt=video) of a DynamicsWorldModel with its tokenizer attached (D4:6377-6387, 6699-6724): tokenize the
    prompt, roll out over the prefilled cache, decode everything; only the video comes back."""
    fx, sd, cfg, tokenizer = _world()
    ref = fx['prompted']
    torch.manual_seed(ref['seed'])
    exp = O.generate(sd, cfg, ref['time_steps'], fx['prompt'].shape[0], tokenizer=tokenizer, prompt=fx['prompt'],
                     return_agent_actions=False, return_decoded_video=True)
    assert exp.video.shape == ref['video'].shape
    torch.testing.assert_close(exp.video, ref['video'], **TOL)


def test_dream_with_decoded_video_matches_reference():
    """The DreamTrainer-flag rollout with return_decoded_video: Experience.video, decoded after the rollout's own draws."""
    fx, sd, cfg, tokenizer = _world()
    ref = fx['dream']
    torch.manual_seed(ref['seed'])
    exp = O.generate(sd, cfg, ref['time_steps'], fx['prompt'].shape[0], tokenizer=tokenizer, return_decoded_video=True)
    assert torch.equal(exp.actions, ref['actions'])
    torch.testing.assert_close(exp.latents, ref['latents'], **TOL)
    torch.testing.assert_close(exp.rewards, ref['rewards'], **TOL)
    torch.testing.assert_close(exp.video, ref['video'], **TOL)


@pytest.mark.parametrize('case', range(4), ids=['vec_mixed_bootstrap', 'vec_all_terminated', 'single_truncated', 'single_terminated'])
def test_interact_with_env_matches_reference(case):
    """interact_with_env (D4:5470-5889) on the deterministic toy image env: per-step incremental tokenize + one clean world-model
    pass + value / policy / sampling, termination and truncation bookkeeping, the bootstrap step and its right-padding."""
    from oracle.toy_env import ToyImageEnv
    fx, sd, cfg, tokenizer = _world()
    ref = fx['interact'][case]
    env = ToyImageEnv(batch=3 if ref['vectorized'] else None, terminate_at=ref['terminate_at'])
    torch.manual_seed(ref['seed'])
    exp = O.interact_with_env(sd, cfg, tokenizer, env, max_timesteps=ref['max_timesteps'], env_is_vectorized=ref['vectorized'])
    assert torch.equal(exp.actions, ref['actions'])
    assert torch.equal(exp.lens, ref['lens']) and torch.equal(exp.is_truncated, ref['is_truncated']) and torch.equal(exp.terminals, ref['terminals'])
    for name in ('latents', 'agent_embed', 'rewards', 'values', 'log_probs', 'old_action_unembeds', 'episode_return'):
        torch.testing.assert_close(getattr(exp, name), ref[name], **TOL, msg=lambda m, n=name: f'{n}: {m}')
    assert ref['video'].shape[2] == exp.rewards.shape[1]            # frames the env showed, bootstrap frame included (D4:5856)
