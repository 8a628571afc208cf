"""Random test environments with the reference's `MockEnv` / `MockDictEnv` interface (reference dreamer4/mocks.py:1-146): image
observations of a fixed shape, uniform rewards, optional random termination / truncation once a step count has passed.  Test
fixtures only - nothing here is on the hot path."""
import torch
from torch import nn


class MockEnv(nn.Module):
    def __init__(self, image_shape, reward_range=(-100, 100), num_envs=1, vectorized=False, terminate_after_step=None,
                 rand_terminate_prob=0.05, can_truncate=False, rand_truncate_prob=0.05):
        super().__init__()
        assert not (vectorized and num_envs == 1)
        self.image_shape, self.reward_range = tuple(image_shape), reward_range
        self.num_envs, self.vectorized = num_envs, vectorized
        self.terminate_after_step, self.rand_terminate_prob = terminate_after_step, rand_terminate_prob
        self.can_terminate = terminate_after_step is not None
        self.can_truncate, self.rand_truncate_prob = can_truncate, rand_truncate_prob
        self.register_buffer('_step', torch.tensor(0))

    def _frame(self):
        frame = torch.randn(3, *self.image_shape)
        return frame.expand(self.num_envs, *frame.shape).clone() if self.vectorized else frame

    def reset(self, seed=None):
        self._step.zero_()
        return self._frame()

    def step(self, actions):
        reward = torch.empty(()).uniform_(*self.reward_range)
        if self.vectorized:
            first = actions[0] if isinstance(actions, tuple) else actions
            assert first.shape[0] == self.num_envs, f'expected batch of actions for {self.num_envs} environments'
            reward = reward.expand(self.num_envs).clone()
        out = (self._frame(), reward)
        if self.can_terminate:
            shape = (self.num_envs,) if self.vectorized else (1,)
            armed = self._step > self.terminate_after_step
            terminate = (torch.rand(shape) < self.rand_terminate_prob) & armed
            out = (*out, terminate)
            if self.can_truncate:
                out = (*out, (torch.rand(shape) < self.rand_truncate_prob) & armed & ~terminate)
        self._step.add_(1)
        return out


class MockDictEnv(nn.Module):
    def __init__(self, image_shape, dim_proprio, num_envs=1, vectorized=False, terminate_after_step=None):
        super().__init__()
        self.image_shape, self.dim_proprio = tuple(image_shape), dim_proprio
        self.num_envs, self.vectorized, self.terminate_after_step = num_envs, vectorized, terminate_after_step
        self.register_buffer('_step', torch.tensor(0))

    def _obs(self):
        lead = (self.num_envs,) if self.vectorized else ()
        return dict(image=torch.randn(*lead, 3, *self.image_shape), proprio=torch.randn(*lead, self.dim_proprio))

    def reset(self):
        self._step.zero_()
        return self._obs()

    def step(self, actions):
        self._step.add_(1)
        reward = torch.randn(self.num_envs) if self.vectorized else torch.randn(())
        done = self.terminate_after_step is not None and bool(self._step >= self.terminate_after_step)
        terminated = torch.full((self.num_envs,), done) if self.vectorized else torch.tensor(done)
        return self._obs(), reward, terminated
