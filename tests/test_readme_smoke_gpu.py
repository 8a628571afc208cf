"""The reference's README walk-through (README.md:37-104; BASELINE.json configs[0]) through the `dreamer4` import name, on the GPU:
tokenizer + world model at the README sizes (256 x 256, patch 32, dim 512), the world-model loss on raw video (forward only here),
`generate(10, batch_size=2, return_decoded_video=True, return_for_policy_optimization=True)`, `learn_from_experience` + backward on
the dream, `interact_with_env` on the vectorized MockEnv + `learn_from_experience` + backward.  What the README also does and this
package does not: `.backward()` through the tokenizer / world-model pre-training losses (outside the hot path)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_readme_walkthrough_runs_on_the_native_path():
    from dreamer4 import VideoTokenizer, DynamicsWorldModel
    from dreamer4.mocks import MockEnv
    torch.manual_seed(0)
    tokenizer = VideoTokenizer(dim=512, dim_latent=32, patch_size=32, image_height=256, image_width=256)
    world_model = DynamicsWorldModel(dim=512, dim_latent=32, video_tokenizer=tokenizer, num_discrete_actions=4).cuda()

    video = torch.randn(2, 3, 10, 256, 256, device='cuda')
    discrete_actions = torch.randint(0, 4, (2, 10, 1), device='cuda')
    rewards = torch.randn(2, 10, device='cuda')
    loss = world_model(video=video, rewards=rewards, discrete_actions=discrete_actions)
    assert loss.ndim == 0 and bool(torch.isfinite(loss))

    dreams = world_model.generate(10, batch_size=2, return_decoded_video=True, return_for_policy_optimization=True)
    frames = dreams.latents.shape[1]
    assert 1 <= frames <= 10 and dreams.video.shape == (2, 3, frames, 256, 256) and bool(torch.isfinite(dreams.video).all())
    actor_loss, critic_loss = world_model.learn_from_experience(dreams)
    (actor_loss + critic_loss).backward()
    grads = [p.grad for p in world_model.policy_head_parameters() + world_model.value_head_parameters() if p.numel() > 0]      # (the empty continuous unembed has none)
    assert all(g is not None and bool(torch.isfinite(g).all()) for g in grads) and any(float(g.abs().sum()) > 0 for g in grads)

    mock_env = MockEnv((256, 256), vectorized=True, num_envs=4)
    experience = world_model.interact_with_env(mock_env, max_timesteps=8, env_is_vectorized=True)
    assert experience.latents.shape[0] == 4 and experience.video is not None
    actor_loss, critic_loss = world_model.learn_from_experience(experience)
    (actor_loss + critic_loss).backward()
    assert bool(torch.isfinite(actor_loss)) and bool(torch.isfinite(critic_loss))
