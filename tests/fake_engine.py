"""A CPU stand-in for libd4b200.so's frame-level entry points, for testing the HOST mirror only (test infrastructure).

`dreamer4_b200.dynamics.DynamicsWorldModel.generate` is bookkeeping around d4_pass / d4_frame: prompts, the action history
that conditions each frame, the in-place KV buffer handed back and forth as `time_cache`, capacity growth, output slicing.
This object implements those C-ABI calls (same argument order as include/d4b200.h) on host memory with the oracle's
arithmetic, so the CPU suite can hold that bookkeeping to `oracle.generate` bit for bit.  It never runs in the product:
tests install it with `install(monkeypatch, model, oracle_cfg)`; the CUDA kernels themselves are covered by the -m gpu tests."""
import contextlib
import ctypes as C
import math

import torch

from oracle import dreamer4_oracle as O


def _addr(p):
    if p is None:
        return 0
    return p.value or 0 if isinstance(p, C.c_void_p) else int(p)


def _flat(p, numel, dtype=torch.float32):
    size = torch.empty(0, dtype=dtype).element_size()
    return torch.frombuffer((C.c_char * (numel * size)).from_address(_addr(p)), dtype=dtype)


def _rows(p, B, row, stride, dtype=torch.float32):
    """(B, row) view of rows `stride` elements apart."""
    return torch.as_strided(_flat(p, (B - 1) * stride + row, dtype), (B, row), (stride, 1))


class FakeEngine:
    def __init__(self, model, ocfg):
        self.model, self.cfg = model, ocfg
        self.kv = None
        self.calls = dict(ctx_create=0, pass_=0, frame=0)

    @property
    def sd(self):
        return {k: v.detach() for k, v in self.model.state_dict().items()}

    # ---- lifetime / plumbing
    def d4_last_error(self):
        return b'fake engine'

    def d4_ctx_create(self, cc, out):
        cc = cc._obj
        self.max_batch, self.max_time = cc.max_batch, cc.max_time
        out._obj.value = 0xD4
        self.calls['ctx_create'] += 1
        return 0

    def d4_ctx_destroy(self, ctx):
        self.kv = None

    def d4_set_weight(self, ctx, name, p, numel):
        return 0

    def d4_bind(self, ctx):
        return 0

    def d4_workspace_bytes(self, ctx):
        return 256

    def _kv_shape(self):
        c = self.cfg
        S = c.tokens_per_frame
        return (max(sum(c.is_time), 1), 2, self.max_batch * S, c.attn_heads, self.max_time, c.attn_dim_head), S

    def d4_kv_bytes(self, ctx):
        return math.prod(self._kv_shape()[0]) * 4

    def d4_set_buffers(self, ctx, ws, ws_bytes, kv, kv_bytes):
        shape, _ = self._kv_shape()
        assert kv_bytes == math.prod(shape) * 4
        self.kv = _flat(kv, kv_bytes // 4).view(shape)
        return 0

    # ---- the hot path, on the oracle's arithmetic
    def _pass(self, B, x, signal, step_log2, prev_actions, tasks, t, commit):
        assert 0 <= t < self.max_time and 1 <= B <= self.max_batch
        L = sum(self.cfg.is_time)
        _, S = self._kv_shape()
        cache = None if t == 0 else [(self.kv[l, 0, :B * S, :, :t], self.kv[l, 1, :B * S, :, :t]) for l in range(L)]
        pred, agent, new_kv = O.forward_step(self.sd, self.cfg, x, signal, step_log2, prev_actions, cache, t, tasks)
        if commit:
            for l in range(L):
                self.kv[l, 0, :B * S, :, t] = new_kv[l][0][:, :, t]
                self.kv[l, 1, :B * S, :, t] = new_kv[l][1][:, :, t]
        return pred, agent

    def _inputs(self, B, prev_actions, pa_stride, tasks):
        na = len(self.cfg.num_discrete_actions)
        pa = _rows(prev_actions, B, na, pa_stride, torch.long).clone() if _addr(prev_actions) else None
        tk = _flat(tasks, B, torch.long).clone() if _addr(tasks) else None
        return pa, tk

    def d4_pass(self, ctx, B, latent, signal, step_log2, prev_actions, pa_stride, tasks, t, commit, pred_out, agent_out, stream):
        c = self.cfg
        x = _flat(latent, B * c.num_latent_tokens * c.dim_latent).view(B, c.num_latent_tokens, c.dim_latent).clone()
        pa, tk = self._inputs(B, prev_actions, pa_stride, tasks)
        pred, agent = self._pass(B, x, signal, step_log2, pa, tk, t, commit)
        if _addr(pred_out):
            _flat(pred_out, pred.numel()).copy_(pred.reshape(-1))
        if _addr(agent_out):
            _flat(agent_out, agent.numel()).copy_(agent.reshape(-1))
        self.calls['pass_'] += 1
        return 0

    def d4_observe(self, ctx, B, t, num_steps, temperature, io, stream):
        self.calls['observe'] = self.calls.get('observe', 0) + 1
        return self.d4_frame(ctx, B, t, num_steps, temperature, io, stream, first_step=num_steps)

    def d4_frame(self, ctx, B, t, num_steps, temperature, io, stream, first_step=0):
        io, c, sd = io._obj, self.cfg, self.sd
        N, Dl, D = c.num_latent_tokens, c.dim_latent, c.dim
        step_size = c.max_steps // num_steps
        step_log2 = int(math.log2(step_size))
        pa, tk = self._inputs(B, io.prev_actions, io.pa_stride, io.tasks)
        x = _flat(io.noise_latent, B * N * Dl).view(B, N, Dl).clone()
        for step in range(first_step, num_steps + 1):
            signal = min(step * step_size, c.max_steps - 1)
            pred, agent = self._pass(B, x, signal, step_log2, pa, tk, t, step == num_steps)
            if step < num_steps:
                x = x + (pred - x) / (1.0 - signal / c.max_steps) * (step_size / c.max_steps)
        _rows(io.latents, B, N * Dl, io.latents_bs).copy_(x.clamp(-1., 1.).reshape(B, -1))
        if _addr(io.agent_embed):
            _rows(io.agent_embed, B, D, io.agent_bs).copy_(agent)
        if _addr(io.rewards):
            codec = O.HLGauss(c.reward_range, c.reward_num_bins, c.hl_gauss_sigma_to_bin_ratio, c.hl_gauss_eps)
            logits = O.rmsnorm(agent, sd['to_reward_pred.nets.0.0.weight']) @ sd['to_reward_pred.nets.0.1.weight'].T
            _rows(io.rewards, B, 1, io.rewards_bs).copy_(codec.from_logits(logits)[:, None])
        if c.predict_terminals and _addr(io.terminal_uniform) and _addr(io.lens) and _addr(io.terminals):
            logit = O.mlp(sd, 'to_state_terminal_pred.0.', x.mean(dim=1), c.head_activation)[..., 0]
            is_term = _flat(io.terminal_uniform, B) < logit.sigmoid()
            lens, term = _flat(io.lens, B, torch.long), _flat(io.terminals, B, torch.uint8)
            lens.masked_fill_(is_term & (term == 0), t + 1)
            term.copy_(((term != 0) | is_term).to(torch.uint8))
        if c.has_actions and _addr(io.actions):
            sizes = list(c.num_discrete_actions)
            logits = O.unembed_logits(sd, O.mlp(sd, 'policy_head.', agent, c.head_activation))
            if _addr(io.logits):
                _rows(io.logits, B, sum(sizes), io.logits_bs).copy_(logits)
            u = _flat(io.action_uniform, B * sum(sizes)).view(B, -1)
            acts, lps = [], []
            for l, uu in zip(logits.split(sizes, dim=-1), u.split(sizes, dim=-1)):
                idx = (l / max(temperature, 1e-10) - O._log(-O._log(uu))).argmax(dim=-1)
                acts.append(idx)
                lps.append(l.log_softmax(dim=-1).gather(-1, idx[:, None])[:, 0])
            _rows(io.actions, B, len(sizes), io.actions_bs, torch.long).copy_(torch.stack(acts, dim=-1))
            _rows(io.log_probs, B, len(sizes), io.log_probs_bs).copy_(torch.stack(lps, dim=-1))
            if _addr(io.values):
                codec = O.HLGauss(c.value_range, c.value_num_bins, c.hl_gauss_sigma_to_bin_ratio, c.hl_gauss_eps)
                _rows(io.values, B, 1, io.values_bs).copy_(codec.from_logits(O.mlp(sd, 'value_head.', agent, c.head_activation))[:, None])
        self.calls['frame'] += 1
        return 0


class _Stream:
    cuda_stream = 0


def install(monkeypatch, model, ocfg):
    """Routes `model`'s native calls to a FakeEngine and lifts the CUDA-only guards (CPU tests of the host logic)."""
    from dreamer4_b200 import _lib
    fake = FakeEngine(model, ocfg)
    monkeypatch.setattr(_lib, 'load', lambda: fake)
    monkeypatch.setattr(type(model), '_require_cuda', lambda self: None)
    monkeypatch.setattr(torch.cuda, 'current_stream', lambda device=None: _Stream())
    monkeypatch.setattr(torch.cuda, 'device', lambda device=None: contextlib.nullcontext())
    return fake
