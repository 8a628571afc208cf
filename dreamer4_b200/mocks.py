"""Random stand-in environments behind the reference's `MockEnv` / `MockDictEnv` interface (reference dreamer4/mocks.py:1-146):
image observations of a fixed shape, random rewards, and - when asked for - random episode ends once a step count has passed.
Test fixtures only; nothing here is on the hot path.  Both classes sit on one small base that owns the step counter and the
"one observation per environment" batching rule, so the two only say what an observation and a transition look like."""
import torch
from torch import nn


class _RandomEnv(nn.Module):
    """Step counter + batching: vectorized environments return a leading `num_envs` axis, single ones do not."""

    def __init__(self, num_envs, vectorized):
        super().__init__()
        if vectorized and num_envs == 1:
            raise AssertionError('a vectorized mock environment needs num_envs > 1')
        self.num_envs, self.vectorized = num_envs, vectorized
        self.register_buffer('_step', torch.tensor(0))

    @property
    def _lead(self):
        return (self.num_envs,) if self.vectorized else ()

    def _flag_shape(self):
        return (self.num_envs,) if self.vectorized else (1,)

    def _restart(self):
        self._step.zero_()


class MockEnv(_RandomEnv):
    """Image observations (3, *image_shape); reward ~ U(reward_range), one draw shared by all environments of a step; after
    `terminate_after_step` steps every environment may terminate (and, with `can_truncate`, otherwise truncate) at random."""

    def __init__(self, image_shape, reward_range=(-100, 100), num_envs=1, vectorized=False, terminate_after_step=None,
                 rand_terminate_prob=0.05, can_truncate=False, rand_truncate_prob=0.05):
        super().__init__(num_envs, vectorized)
        self.image_shape = tuple(image_shape)
        self.reward_range = reward_range
        self.terminate_after_step = terminate_after_step
        self.can_terminate = terminate_after_step is not None
        self.rand_terminate_prob = rand_terminate_prob
        self.can_truncate = can_truncate
        self.rand_truncate_prob = rand_truncate_prob

    def _observe(self):
        image = torch.randn(3, *self.image_shape)                       # one random frame, repeated for every environment
        return image.repeat(self.num_envs, 1, 1, 1) if self.vectorized else image

    def reset(self, seed=None):
        self._restart()
        return self._observe()

    def _episode_flags(self):
        """(terminate,) or (terminate, truncate): random, and only once the step counter has passed the threshold."""
        past = self._step > self.terminate_after_step
        terminate = past & (torch.rand(self._flag_shape()) < self.rand_terminate_prob)
        if not self.can_truncate:
            return (terminate,)
        truncate = past & ~terminate & (torch.rand(self._flag_shape()) < self.rand_truncate_prob)
        return (terminate, truncate)

    def step(self, actions):
        if self.vectorized:
            batch = (actions[0] if isinstance(actions, tuple) else actions).shape[0]
            assert batch == self.num_envs, f'expected batch of actions for {self.num_envs} environments'
        low, high = self.reward_range
        reward = torch.empty(()).uniform_(low, high)
        if self.vectorized:
            reward = reward.repeat(self.num_envs)
        observation = self._observe()
        flags = self._episode_flags() if self.can_terminate else ()
        self._step.add_(1)
        return (observation, reward, *flags)


class MockDictEnv(_RandomEnv):
    """Observations are dicts {image (3, *image_shape), proprio (dim_proprio,)}; normal rewards; every environment terminates
    together once `terminate_after_step` steps have been taken."""

    def __init__(self, image_shape, dim_proprio, num_envs=1, vectorized=False, terminate_after_step=None):
        super().__init__(num_envs, vectorized)
        self.image_shape, self.dim_proprio = tuple(image_shape), dim_proprio
        self.terminate_after_step = terminate_after_step

    def _observe(self):
        return dict(image=torch.randn(*self._lead, 3, *self.image_shape), proprio=torch.randn(*self._lead, self.dim_proprio))

    def reset(self):
        self._restart()
        return self._observe()

    def step(self, actions):
        self._step.add_(1)
        over = self.terminate_after_step is not None and int(self._step) >= self.terminate_after_step
        reward = torch.randn(self._lead)
        terminated = torch.full(self._lead, over, dtype=torch.bool)
        return self._observe(), reward, terminated
