"""Rollout record exchanged between `generate` and `learn_from_experience`.

Field names, dtypes and shapes follow the reference's `Experience` / `Actions`
(reference dreamer4/dreamer4.py:132-309); the replay-buffer (de)hydration helpers of the reference are
storage code outside this path and are not provided."""
from __future__ import annotations

from collections import namedtuple
from dataclasses import dataclass, fields
from typing import Optional

import torch
from torch import Tensor

Actions = namedtuple('Actions', ['discrete', 'continuous'])

TransformerIntermediates = namedtuple('TransformerIntermediates', ['next_kv_cache', 'token_count'])
DynamicsIntermediates = namedtuple('DynamicsIntermediates', ['main'])
Predictions = namedtuple('Predictions', ['flow', 'proprioception', 'state'])              # reference dreamer4.py:128
Embeds = namedtuple('Embeds', ['agent', 'state_pred', 'actor', 'critic'], defaults=(None, None, None))    # reference dreamer4.py:130


def _map_tensors(fn, v):
    if torch.is_tensor(v):
        return fn(v)
    if isinstance(v, Actions):
        return Actions(*(_map_tensors(fn, x) for x in v))
    return v


@dataclass
class Experience:
    latents: Tensor
    video: Optional[Tensor] = None
    proprio: Optional[Tensor] = None
    critic_state: Optional[Tensor] = None
    agent_embed: Optional[Tensor] = None
    rewards: Optional[Tensor] = None
    terminals: Optional[Tensor] = None
    actions: Optional[Actions] = None
    log_probs: Optional[Actions] = None
    old_action_unembeds: Optional[Actions] = None
    values: Optional[Tensor] = None
    step_size: Optional[int] = None
    lens: Optional[Tensor] = None
    is_truncated: Optional[Tensor] = None
    agent_index: int = 0
    is_from_world_model: bool = True
    episode_return: Optional[Tensor] = None

    def to(self, device):
        return Experience(**{f.name: _map_tensors(lambda t: t.to(device), getattr(self, f.name)) for f in fields(self)})

    def cpu(self):
        return self.to(torch.device('cpu'))


def combine_experiences(exps):
    """Concatenates experiences along the batch dimension, right-padding time to the longest
    (reference dreamer4/dreamer4.py:248-309)."""
    assert len(exps) > 0
    max_t = max(e.latents.shape[1] for e in exps)

    def pad_t(t):
        if t is None or t.ndim < 2 or t.shape[1] == max_t:
            return t
        pad = [0, 0] * (t.ndim - 2) + [0, max_t - t.shape[1]]
        return torch.nn.functional.pad(t, pad)

    def cat(vals, time_dim=True):
        if any(v is None for v in vals):
            return None
        if isinstance(vals[0], Actions):
            return Actions(cat([v.discrete for v in vals]), cat([v.continuous for v in vals]))
        if not torch.is_tensor(vals[0]):
            return vals[0]
        return torch.cat([pad_t(v) if time_dim else v for v in vals], dim=0)

    out = {}
    batch_only = {'lens', 'is_truncated', 'terminals', 'episode_return'}
    for f in fields(Experience):
        vals = [getattr(e, f.name) for e in exps]
        if f.name == 'lens':
            vals = [v if v is not None else torch.full((e.latents.shape[0],), e.latents.shape[1], device=e.latents.device)
                    for v, e in zip(vals, exps)]
        out[f.name] = cat(vals, time_dim=f.name not in batch_only)
    return Experience(**out)
