"""`DreamTrainer` — training the policy / value heads from imagined rollouts (reference dreamer4/trainers.py:1330-1468).

Same constructor keywords and loop as the reference; HF `accelerate` (the reference's only distributed layer) is replaced
by one process per GPU over `torch.distributed`: dreams are sharded across ranks, and each update does one flat gradient
all-reduce (dreamer4_b200/dist.py)."""
from __future__ import annotations

import torch
from torch import nn
from torch.optim import AdamW

from . import dist as D
from .dynamics import DynamicsWorldModel, exists


class DreamTrainer(nn.Module):
    def __init__(self, model: DynamicsWorldModel, optim_klass=AdamW, batch_size=16, generate_timesteps=16, learning_rate=3e-4,
                 max_grad_norm=0.5, num_train_steps=10_000, weight_decay=0., objective='ppo', optim_kwargs: dict = dict(),
                 cpu=False, **ignored_logging_kwargs):
        super().__init__()
        if cpu:
            raise NotImplementedError('cpu=True: the B200 hot path has no CPU fallback')
        self.model = model
        self.objective = objective
        kw = dict(lr=learning_rate, weight_decay=weight_decay)
        self.policy_head_optim = optim_klass(model.policy_head_parameters(), **kw)
        self.value_head_optim = optim_klass(model.value_head_parameters(), **kw)
        self.max_grad_norm = max_grad_norm
        self.num_train_steps = num_train_steps
        self.batch_size = batch_size                      # per process, as under accelerate
        self.generate_timesteps = generate_timesteps
        self.register_buffer('step', torch.tensor(0))

    @property
    def device(self):
        return self.model.device

    @property
    def unwrapped_model(self):
        return self.model

    @property
    def is_main_process(self):
        return D.rank() == 0

    def print(self, *args, **kwargs):
        if self.is_main_process:
            print(*args, **kwargs)

    def train_step(self, noise=None):
        dreams = self.model.generate(self.generate_timesteps + 1, batch_size=self.batch_size, return_rewards_per_frame=True,
                                     return_agent_actions=True, return_log_probs_and_values=True, noise=noise)     # :1422-1428
        policy_loss, value_loss = self.model.learn_from_experience(dreams, objective=self.objective)             # :1430
        policy_loss.backward()
        value_loss.backward()
        D.allreduce_mean_grads_(self.model.policy_head_parameters() + self.model.value_head_parameters())
        if exists(self.max_grad_norm):
            nn.utils.clip_grad_norm_(self.model.policy_head_parameters(), self.max_grad_norm)                    # :1438-1439
        self.policy_head_optim.step()
        self.policy_head_optim.zero_grad()
        if exists(self.max_grad_norm):
            nn.utils.clip_grad_norm_(self.model.value_head_parameters(), self.max_grad_norm)                     # :1448-1449
        self.value_head_optim.step()
        self.value_head_optim.zero_grad()
        self.step += 1
        return policy_loss.detach(), value_loss.detach(), dreams

    def forward(self):
        for _ in range(self.num_train_steps):
            pl, vl, _ = self.train_step()
            self.print(f'policy head loss: {pl.item():.3f} | value head loss: {vl.item():.3f}')
        self.print('training complete')
