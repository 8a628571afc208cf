"""Gym-style stepping of a `DynamicsWorldModel` (reference dreamer4/env.py:353-553, `DynamicsWorldModelWrapper`).

Each `step` is one prompted `generate` call for a single new frame over the time cache of everything imagined so far
(reference env.py:464-484) - on this path the cache is the engine's in-place KV buffer, so a step costs the five passes
of one frame and no copy.  With a video tokenizer attached to the world model the observation is the newest decoded frame
`(b, c, h, w)` and a video `prompt` seeds the episode, as in the reference (env.py:405-417, 441, 505: the whole rollout so far is
decoded on every step there, and so it is here); without one it is the newest frame's latent `(b, n, d)`, which is what the
reference returns when `gen_out.video` is absent."""
from __future__ import annotations

import torch

from .dynamics import DynamicsWorldModel, exists


class DynamicsWorldModelWrapper:
    def __init__(self, world_model: DynamicsWorldModel, prompt=None, num_generation_steps=4, max_steps=1000, device=None, image_size=None):
        tokenizer = getattr(world_model, 'video_tokenizer', None)
        if exists(prompt):
            assert exists(tokenizer), 'a video prompt needs a video_tokenizer attached to the world model'
        self.prompt = prompt
        self.decodes = exists(tokenizer)
        image_size = image_size if exists(image_size) else getattr(tokenizer, 'image_height', None)      # env.py:371-372
        self.world_model = world_model.eval()
        self.num_generation_steps = num_generation_steps
        self.max_steps = max_steps
        self.device = device if exists(device) else world_model.device
        self.image_size = image_size
        self.clear()

    def clear(self):
        self._latents = self._discrete_actions = self._rewards = self._time_cache = None
        self._current_step = 0

    # the flags env.py:405-417 / 464-484 pass on every call (decoded video excepted, see the module docstring)
    _FLAGS = dict(return_rewards_per_frame=True, return_terminals=True, use_time_cache=True, return_time_cache=True)

    def reset(self, batch_size=None, seed=None):
        batch_size = batch_size if exists(batch_size) else 1
        if exists(seed):
            torch.manual_seed(seed)
        self.clear()
        if exists(self.prompt):
            # The prompt's frames open the episode and the first imagined frame follows them.  (The reference passes the prompt with
            # time_steps=1, env.py:405-417, which returns the prompt alone and then fails forming episode_return - rewards (b, 0)
            # against a (b, P) step mask, dreamer4.py:6741-6743 - so there is no reference behaviour to match here; prompt frames
            # carry zero reward.)
            prompt = self.prompt if self.prompt.ndim == 5 else self.prompt[:, :, None]
            batch_size, frames = prompt.shape[0], prompt.shape[2] + 1
            gen, self._time_cache = self.world_model.generate(time_steps=frames, batch_size=batch_size, num_steps=self.num_generation_steps, prompt=prompt,
                                                              prompt_rewards=torch.zeros(batch_size, frames - 1, device=self.device), **self._flags())
        else:
            gen, self._time_cache = self.world_model.generate(time_steps=1, batch_size=batch_size, num_steps=self.num_generation_steps, **self._flags())
        self._latents, self._rewards = gen.latents, gen.rewards
        na = len(self.world_model.cfg.num_discrete_actions)
        if na > 0:
            self._discrete_actions = torch.empty(batch_size, 0, na, dtype=torch.long, device=self.device)
        return self._obs(gen), dict()

    def _flags(self):
        return dict(self._FLAGS, return_decoded_video=self.decodes, image_height=self.image_size, image_width=self.image_size)

    @staticmethod
    def _obs(gen):                                             # env.py:441, 505
        return gen.video[:, :, -1] if exists(gen.video) else gen.latents[:, -1]

    def step(self, action):
        assert exists(self._latents), 'call reset() first'
        self._current_step += 1
        action = self._parse_action(action)
        if exists(action) and exists(self._discrete_actions):
            self._discrete_actions = torch.cat((self._discrete_actions, action), dim=1)
        batch, frames = self._latents.shape[:2]
        gen, self._time_cache = self.world_model.generate(
            time_steps=frames + 1, batch_size=batch, num_steps=self.num_generation_steps, prompt_latents=self._latents,
            prompt_discrete_actions=self._discrete_actions, prompt_rewards=self._rewards, time_cache=self._time_cache, **self._flags())
        self._latents, self._rewards = gen.latents, gen.rewards
        reward = gen.rewards[:, -1]
        terminated = gen.terminals if exists(gen.terminals) else torch.zeros(batch, dtype=torch.bool, device=self.device)
        truncated = torch.full((batch,), self._current_step >= self.max_steps, dtype=torch.bool, device=self.device)
        return self._obs(gen), reward, terminated, truncated, dict(experience=gen)

    def _parse_action(self, action):
        """Unbatched or batched discrete actions -> (b, 1, na) int64 (env.py:511-553): a 1-D input is one action vector when
        the env holds a single episode and one scalar action per episode otherwise; a 2-D input is (b, na)."""
        if not exists(action) or len(self.world_model.cfg.num_discrete_actions) == 0:
            return None
        action = torch.atleast_1d(torch.as_tensor(action, device=self.device)).long()
        batch = self._latents.shape[0]
        if action.ndim == 1:
            action = action[None, None, :] if batch == 1 else action[:, None, None]
        elif action.ndim == 2:
            action = action[:, None, :]
        else:
            raise ValueError(f'action must be 1D or 2D, but got {action.ndim}D')
        assert action.shape[0] == batch, f'action batch size {action.shape[0]} must match environment batch size {batch}'
        return action
