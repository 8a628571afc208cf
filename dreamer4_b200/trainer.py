"""`DreamTrainer` — training the policy / value heads from imagined rollouts (reference dreamer4/trainers.py:1330-1468) — and
`SimTrainer`, the same from real-environment episodes (:1472-1790).

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


class SimTrainer(nn.Module):
    """Online RL against a real environment (reference dreamer4/trainers.py:1472-1790): episodes are collected with
    `interact_with_env`, combined, and replayed in shuffled minibatches through `learn_from_experience` for `epochs` passes,
    policy head and value head stepping on their own optimizers.  Same constructor keywords, same minibatch pipeline (a
    TensorDataset behind a shuffling DataLoader, so a seeded run visits the same samples in the same order as the reference's);
    `accelerate` is replaced by the one flat gradient all-reduce of dreamer4_b200/dist.py."""

    def __init__(self, model: DynamicsWorldModel, optim_klass=AdamW, batch_size=16, generate_timesteps=16, learning_rate=3e-4,
                 max_grad_norm=None, epochs=2, weight_decay=0., objective='ppo', optim_kwargs: dict = dict(), cpu=False,
                 **ignored_logging_kwargs):
        super().__init__()
        if cpu:
            raise NotImplementedError('cpu=True: the B200 hot path has no CPU fallback')
        self.model = model
        self.objective = objective
        kw = dict(lr=learning_rate, weight_decay=weight_decay)            # the reference overwrites optim_kwargs with these two (:1513-1516)
        self.policy_head_optim = optim_klass(model.policy_head_parameters(), **kw)
        self.value_head_optim = optim_klass(model.value_head_parameters(), **kw)
        self.max_grad_norm = max_grad_norm
        self.epochs = epochs
        self.batch_size = batch_size
        self.generate_timesteps = generate_timesteps
        self.register_buffer('step', torch.tensor(0))

    device = DreamTrainer.device
    unwrapped_model = DreamTrainer.unwrapped_model
    is_main_process = DreamTrainer.is_main_process
    print = DreamTrainer.print

    def learn(self, experience):
        """reference trainers.py:1559-1696: the per-sample fields of the combined experience, minibatched; lens / truncation flags
        do not travel (every step of a minibatch row is learned from, as in the reference)."""
        from torch.utils.data import DataLoader, TensorDataset
        from .experience import Actions, Experience
        dev = self.device
        rewards = experience.rewards
        empty = torch.empty_like(rewards)
        has_agent_embed = exists(experience.agent_embed)
        old_unembeds = experience.old_action_unembeds.discrete if exists(experience.old_action_unembeds) else None
        has_unembeds = exists(old_unembeds)
        dataset = TensorDataset(experience.latents, experience.actions.discrete, experience.log_probs.discrete,
                                experience.agent_embed if has_agent_embed else empty, old_unembeds if has_unembeds else empty,
                                experience.values, rewards)
        losses = []
        for _ in range(self.epochs):
            for latents, actions, log_probs, agent_embed, unembeds, old_values, rew in DataLoader(dataset, batch_size=self.batch_size, shuffle=True):
                to = lambda t: t.to(dev)
                batch = Experience(latents=to(latents), actions=Actions(to(actions), None), log_probs=Actions(to(log_probs), None),
                                   agent_embed=to(agent_embed) if has_agent_embed else None,
                                   old_action_unembeds=Actions(to(unembeds), None) if has_unembeds else None,
                                   values=to(old_values), rewards=to(rew), step_size=experience.step_size, agent_index=experience.agent_index)
                policy_loss, value_loss = self.model.learn_from_experience(batch, objective=self.objective)
                self.print(f'policy head loss: {policy_loss.item():.3f} | value head loss: {value_loss.item():.3f}')
                losses.append((policy_loss.detach(), value_loss.detach()))
                self.step += 1
                policy_loss.backward()
                value_loss.backward()
                D.allreduce_mean_grads_(self.model.policy_head_parameters() + self.model.value_head_parameters())
                if exists(self.max_grad_norm):
                    nn.utils.clip_grad_norm_(self.model.policy_head_parameters(), self.max_grad_norm)
                self.policy_head_optim.step()
                self.policy_head_optim.zero_grad()
                if exists(self.max_grad_norm):
                    nn.utils.clip_grad_norm_(self.model.value_head_parameters(), self.max_grad_norm)
                self.value_head_optim.step()
                self.value_head_optim.zero_grad()
        return losses

    def forward(self, env, num_episodes=50000, max_experiences_before_learn=8, env_is_vectorized=False, **interact_kwargs):
        """reference trainers.py:1698-1790.  `interact_kwargs` (e.g. obs_to_latents_fn, max_timesteps) are handed to interact_with_env."""
        from .experience import combine_experiences
        for _ in range(num_episodes):
            total, experiences = 0, []
            while total < max_experiences_before_learn:
                experience = self.model.interact_with_env(env, env_is_vectorized=env_is_vectorized, **interact_kwargs)
                total += experience.latents.shape[0]
                experiences.append(experience.cpu())
            self.learn(combine_experiences(experiences))
        self.print('training complete')
