"""A deterministic image environment for the interact_with_env golden vectors (test infrastructure).

Observations are (c, h, w) images that depend on the step counter and on the previous action, rewards depend on the action,
and each episode of the vectorized variant terminates at its own step (or never, so that max_timesteps truncates it and the
bootstrap branch of interact_with_env runs, reference dreamer4.py:5790-5853)."""
import numpy as np
import torch


class ToyImageEnv:
    def __init__(self, batch=None, channels=3, size=16, terminate_at=None):
        """batch=None: a single non-vectorized env (obs (c h w), scalar reward / flags); otherwise vectorized over `batch`.
        terminate_at: per-episode step index at which `terminated` turns True (None or 0 = never)."""
        self.batch, self.channels, self.size = batch, channels, size
        n = 1 if batch is None else batch
        self.terminate_at = torch.as_tensor(terminate_at if terminate_at is not None else [0] * n).reshape(n)
        self.t = 0
        self.actions_seen = []

    def _obs(self, last_action):
        n = self.terminate_at.shape[0]
        g = torch.Generator().manual_seed(1000 + self.t)
        base = torch.rand(n, self.channels, self.size, self.size, generator=g)
        shift = (last_action.reshape(n, -1).sum(dim=-1).float() * 0.05)[:, None, None, None]
        obs = (base + shift).clamp(0., 1.)
        return obs if self.batch is not None else obs[0]

    def reset(self, seed=None):
        self.t = 0
        self.actions_seen = []
        return self._obs(torch.zeros(self.terminate_at.shape[0], 1)), dict()

    def step(self, action):
        self.t += 1
        n = self.terminate_at.shape[0]
        a = torch.as_tensor(np.asarray(action)).reshape(n, -1)
        self.actions_seen.append(a.clone())
        reward = a.sum(dim=-1).float() * 0.5 - 0.25 * self.t
        terminated = (self.terminate_at > 0) & (self.t >= self.terminate_at)
        truncated = torch.zeros(n, dtype=torch.bool)
        obs = self._obs(a)
        if self.batch is None:
            return obs, float(reward[0]), bool(terminated[0]), bool(truncated[0]), dict()
        return obs, reward.numpy(), terminated.numpy(), truncated.numpy(), dict()
