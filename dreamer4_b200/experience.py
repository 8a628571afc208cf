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
WorldModelLosses = namedtuple('WorldModelLosses', ('flow', 'shortcut', 'rewards', 'terminals', 'discrete_actions', 'continuous_actions', 'state_pred',
                                                    'agent_state_pred', 'latent_ar', 'latent_ar_sigreg', 'lapo_action', 'lapo_fdm', 'lapo_raw_latent_fdm', 'tem',
                                                    'h_net'))          # reference dreamer4.py:120
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


def _pad_to(t, dim, size):
    if t.ndim <= dim or t.shape[dim] == size:
        return t
    pad = [0, 0] * (t.ndim - 1 - dim) + [0, size - t.shape[dim]]
    return torch.nn.functional.pad(t, pad)


def combine_experiences(exps):
    """Concatenates experiences along the batch dimension (reference dreamer4/dreamer4.py:248-309).  Every tensor field is
    right-padded with zeros along dims 1 and 2 to the longest among the inputs before the concatenation - dim 1 is time for
    (b, t, ...) fields and dim 2 is time for `video` (b, c, t, h, w), the reference's `pad_tensors_at_dim_to_max_len(dims = (1, 2))`;
    missing `lens` default to the full length, missing `is_truncated` to True, a bool `is_from_world_model` becomes a (b,) tensor.
    Non-tensor fields (step_size, agent_index) must agree."""
    assert len(exps) > 0
    for e in exps:
        payload = e.latents if e.latents is not None else e.video
        b, t, dev = payload.shape[0], payload.shape[1], payload.device
        if e.lens is None:
            e.lens = torch.full((b,), t, device=dev)
        if e.is_truncated is None:
            e.is_truncated = torch.full((b,), True, device=dev)
        if isinstance(e.is_from_world_model, bool):
            e.is_from_world_model = torch.full((b,), e.is_from_world_model, device=dev, dtype=torch.bool)

    def join(vals, name):
        if isinstance(vals[0], Actions):
            assert all(isinstance(v, Actions) for v in vals), f'{name}: some experiences carry it, some do not'
            return Actions(join([v.discrete for v in vals], name + '.discrete'), join([v.continuous for v in vals], name + '.continuous'))
        if torch.is_tensor(vals[0]):
            assert all(torch.is_tensor(v) for v in vals), f'{name}: some experiences carry it, some do not'
            for dim in (1, 2):
                size = max((v.shape[dim] for v in vals if v.ndim > dim), default=0)
                vals = [_pad_to(v, dim, size) for v in vals]
            return torch.cat(vals) if vals[0].ndim > 0 else torch.stack(vals)
        assert all(v == vals[0] for v in vals), f'{name}: experiences disagree ({vals})'
        return vals[0]

    return Experience(**{f.name: join([getattr(e, f.name) for e in exps], f.name) for f in fields(Experience)})
