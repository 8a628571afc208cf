"""Host-side mirror of the reference's `DynamicsWorldModel` for the imagination hot path.

Same constructor keywords, parameter names (state_dict keys) and method signatures as the reference
(reference dreamer4/dreamer4.py:4660-4778 constructor, 6307-6774 `generate`, 5893-6305 `learn_from_experience`),
with the per-frame work handed to the hand-written sm_100a kernels behind include/d4b200.h.  PyTorch is used for
device memory, the RNG stream and autograd plumbing only.  There is no CPU / eager fallback: the model must live on
a CUDA device and libd4b200.so must be built, otherwise the calls raise."""
from __future__ import annotations

import ctypes as C
import functools
import math
import pickle
from dataclasses import dataclass
from typing import Optional

import torch
from torch import nn

from . import _lib
from ._lib import D4Error, check, ptr
from .experience import Actions, DynamicsIntermediates, Embeds, Experience, Predictions, TransformerIntermediates
from . import registry
from .packing import hl_gauss_tables, mlp_param_names, pack, tf32_split


def exists(v):
    return v is not None


def default(v, d):
    return v if exists(v) else d


@dataclass
class ModelConfig:
    dim: int
    dim_latent: int
    num_latent_tokens: int
    num_spatial_tokens: int = 4
    num_register_tokens: int = 8
    depth: int = 4
    time_block_every: int = 4
    attn_heads: int = 8
    query_heads: int = 8
    attn_dim_head: int = 64
    attn_softclamp_value: float = 50.
    max_steps: int = 64
    num_discrete_actions: tuple = ()
    multi_token_pred_len: int = 8
    policy_head_mlp_depth: int = 3
    value_head_mlp_depth: int = 3
    terminal_mlp_depth: int = 1
    ff_activation: str = 'silu'
    ff_expansion_factor: float = 4.
    reward_range: tuple = (-20., 20.)
    reward_num_bins: int = 255
    value_range: tuple = (-20., 20.)
    value_num_bins: int = 255
    hl_gauss_sigma_to_bin_ratio: float = 2.
    hl_gauss_eps: float = 1e-10
    predict_terminals: bool = True
    num_tasks: int = 0
    num_agents: int = 1
    pool_heads: int = 4
    pool_dim_head: int = 64

    @property
    def has_actions(self):
        return len(self.num_discrete_actions) > 0

    @property
    def same_len(self):
        return self.num_spatial_tokens == self.num_latent_tokens

    @property
    def tokens_per_frame(self):        # reference dreamer4.py:7222
        return 1 + self.num_spatial_tokens + self.num_register_tokens + int(self.has_actions) + 1

    @property
    def is_time(self):                 # reference dreamer4.py:2845
        return [((i + 1) % self.time_block_every) == 0 for i in range(self.depth)]

    @property
    def num_time_layers(self):
        return sum(self.is_time)

    @property
    def ff_inner(self):                # reference dreamer4.py:2094
        return int(self.dim * self.ff_expansion_factor * 2 / 3)

    @property
    def ff_inner_pad(self):
        return (self.ff_inner + 31) // 32 * 32

    @property
    def total_actions(self):
        return sum(self.num_discrete_actions)


class _Node(nn.Module):
    """Bare container used to reproduce the reference's state_dict key hierarchy."""


class _HeadNode(_Node):
    """`model.policy_head` / `model.value_head` / the terminal head as callables (the reference exposes them as nn.Modules and its
    tests call them directly, e.g. `dynamics.policy_head(embeds.agent)`, tests/test_dreamer.py:1262): the MLP runs natively
    (d4_head_forward); inference only - gradients of the heads come from learn_from_experience."""

    def forward(self, x):
        model, which = self._d4_model(), self._d4_which
        assert model is not None, 'the owning DynamicsWorldModel is gone'
        return model._head_forward(which, x)


def _linear_w(out_f, in_f):
    w = torch.empty(out_f, in_f)
    nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    return w


def _linear_b(out_f, in_f):
    bound = 1 / math.sqrt(in_f) if in_f > 0 else 0
    return torch.empty(out_f).uniform_(-bound, bound)


# Constructor keywords of the reference (dreamer4.py:4662-4778) that select a branch outside the B200 hot path: accepted at their
# default value only.
_UNSUPPORTED_DEFAULTS = dict(
    aux_image_encoder=None, num_video_views=1, mot_temporal=False, dim_proprio=None, dim_state=None,
    dim_critic_state=None, reward_encoder_type='hl_gauss', critic_state_embedder=None, spatial_pre_encoder_depth=0,
    action_pre_encoder_depth=0, actor_depth=0, critic_depth=0, pred_orig_latent=True, use_time_rnn=False,
    add_reward_embed_to_agent_token=False, add_state_pred_head=False, agent_predicts_state=False, num_continuous_actions=0,
    num_latent_genes=0, time_attention_use_pope=False, latent_ar=False, identity_latents_to_spatial=False,
    has_aug_conditioning=False, ssl_lapo=False, ssl_tem=False, actor_spr=False, clip_values=False,
    policy_head_mlp_activation='silu', value_head_mlp_activation='silu', state_terminal_pred_mlp_activation='silu',
)

# Keywords that only steer the world-model / tokenizer TRAINING losses, dropout or optional heads that are off (reference
# dreamer4.py:4662-4778): they never enter generate / interact_with_env / learn_from_experience with the heads detached, so they are
# accepted and recorded but not used.  (`agent_policy_gradient_frac` / `agent_value_gradient_frac` scale the gradient that flows
# back into the agent embedding, which `only_learn_policy_value_heads=True` detaches: no effect on this path.)  Anything that is
# in neither table is a TypeError, as it would be for the reference's explicit signature.
_TRAINING_ONLY = frozenset((
    'freeze_aux_image_encoder', 'loss_weight_fn', 'prob_shortcut_train', 'add_reward_embed_dropout', 'state_pred_loss_weight',
    'state_entropy_bonus_weight', 'agent_predicts_state_frac_gradient', 'agent_state_pred_loss_weight', 'eps_latent_pred',
    'continuous_norm_stats', 'continuous_dist_type', 'continuous_dist_kwargs', 'continuous_target_action_range',
    'latent_flow_loss_weight', 'shortcut_loss_weight', 'reward_loss_weight', 'terminal_loss_weight', 'discrete_action_loss_weight',
    'continuous_action_loss_weight', 'value_clip', 'agent_policy_gradient_frac', 'agent_value_gradient_frac', 'use_loss_normalization',
    'latent_ar_layer', 'latent_ar_action_conditioned', 'latent_ar_loss_weight', 'latent_ar_sigreg_loss_weight',
    'latent_ar_sigreg_loss_kwargs', 'latent_ar_sigreg_num_subspaces', 'latent_ar_kwargs', 'aug_cfg_dropout_prob',
    'agent_predict_sem_kwargs', 'lapo_pred_actions', 'lapo_use_fdm', 'tem_first_state_as_init_hidden', 'tem_learn_relative_actions',
    'lapo_kwargs', 'tem_kwargs', 'lapo_action_loss_weight', 'lapo_fdm_loss_weight', 'lapo_raw_latent_fdm_loss_weight', 'tem_loss_weight',
    'actor_nlp_kwargs', 'agent_state_pred_mlp_activation',
))


def _records_config(init):
    """Keeps the constructor arguments on the instance, as `torch_einops_utils.save_load` does for the reference class
    (reference dreamer4.py:72, 4660) - what .save / .init_and_load pickle next to the state_dict."""
    @functools.wraps(init)
    def wrapped(self, *args, **kwargs):
        self._config = (args, kwargs)
        init(self, *args, **kwargs)
    return wrapped


class DynamicsWorldModel(nn.Module):
    """Drop-in for the reference class on the imagination path (generate + learn_from_experience).

    Extra keyword arguments (not in the reference): `precision` in {'tf32x3' (default), 'fp32', 'tf32', 'f16x3' (experimental)} — arithmetic of the
    dense layers: 3-term TF32 split on the tensor cores (fp32-accurate), exact-fp32 FMA, or single-pass TF32 (reduced
    precision); the final action unembedding always runs exact fp32.  `time_attn_variant`: K1 kernel (1 = bulk-copy ring, 0 = ld.global)."""

    @_records_config
    def __init__(self, dim, dim_latent, video_tokenizer=None, copy_video_tokenizer=True, *, num_latent_tokens=None, max_steps=64,
                 num_register_tokens=8, num_spatial_tokens=4,
                 num_agents=1, num_tasks=0, reward_encoder_kwargs: dict = dict(), value_encoder_kwargs: Optional[dict] = None,
                 depth=4, time_block_every=4, attn_kwargs: dict = dict(), transformer_kwargs: dict = dict(), attn_heads=8,
                 attn_dim_head=64, attn_softclamp_value=50., ff_kwargs: dict = dict(), num_discrete_actions=0,
                 multi_token_pred_len=8, value_head_mlp_depth=3, policy_head_mlp_depth=3, predict_terminals=True,
                 predict_terminal_mlp_kwargs: dict = dict(depth=1), keep_reward_ema_stats=False, reward_ema_decay=0.998,
                 reward_quantile_filter=(0.05, 0.95), gae_discount_factor=0.997, gae_lambda=0.95, ppo_eps_clip=0.2,
                 use_delight_gating=True, delight_temperature=1., normalize_advantages=None, policy_entropy_weight=.01,
                 pmpo_pos_to_neg_weight=0.5, pmpo_reverse_kl=True, pmpo_kl_div_loss_weight=.3, gae_use_accelerated=False, precision='tf32x3', time_attn_variant=1, **kwargs):
        super().__init__()
        # `video_tokenizer` is the reference's third positional argument (dreamer4.py:4666); `copy_video_tokenizer` deep-copies and
        # freezes it (4789-4792)
        if exists(video_tokenizer):             # reference dreamer4.py:4788-4801
            assert hasattr(video_tokenizer, 'tokenize') and hasattr(video_tokenizer, 'decode'), 'video_tokenizer must be a dreamer4_b200.VideoTokenizer'
            if copy_video_tokenizer:
                import copy
                video_tokenizer = copy.deepcopy(video_tokenizer)
                video_tokenizer.requires_grad_(False)
            num_latent_tokens = default(num_latent_tokens, video_tokenizer.num_latent_tokens)
            assert video_tokenizer.num_latent_tokens == num_latent_tokens and video_tokenizer.dim_latent == dim_latent, \
                'the tokenizer and the dynamics model disagree on the latent shape'
        for k, v in kwargs.items():
            if k in _TRAINING_ONLY:
                continue                    # recorded in self._config; never read on this path (see _TRAINING_ONLY)
            if k not in _UNSUPPORTED_DEFAULTS:
                raise TypeError(f'DynamicsWorldModel.__init__() got an unexpected keyword argument {k!r}')
            if v != _UNSUPPORTED_DEFAULTS[k]:
                raise NotImplementedError(f'{k}={v!r}: this branch of the reference is outside the B200 hot path (SURVEY.md section 8)')
        assert exists(num_latent_tokens), '`num_latent_tokens` must be set (or attach a video_tokenizer)'
        assert precision in _lib.PREC, f'precision must be one of {list(_lib.PREC)}'
        if transformer_kwargs:
            raise NotImplementedError('transformer_kwargs overrides are outside the B200 hot path')
        nda = num_discrete_actions
        nda = (nda,) if isinstance(nda, int) else tuple(nda)
        nda = tuple(int(n) for n in nda if n > 0)
        rk = dict(reward_encoder_kwargs or {})
        vk = rk if value_encoder_kwargs is None else dict(value_encoder_kwargs)
        ffk = dict(ff_kwargs or {})
        act = ffk.get('activation', 'silu')
        if act not in registry.NATIVE_FF_ACTIVATIONS:       # includes names added through register_activation (reference :568-569)
            assert isinstance(act, str) and act in registry.ACTIVATIONS, f'activation {act} not found in {list(registry.ACTIVATIONS.keys())}'
            raise NotImplementedError(f"feed-forward activation {act!r}: the GEMM epilogues cover the gated silu / gelu variants")
        self.cfg = cfg = ModelConfig(
            dim=dim, dim_latent=dim_latent, num_latent_tokens=num_latent_tokens, num_spatial_tokens=num_spatial_tokens,
            num_register_tokens=num_register_tokens, depth=depth, time_block_every=time_block_every, attn_heads=attn_heads,
            query_heads=(attn_kwargs or {}).get('query_heads') or attn_heads, attn_dim_head=attn_dim_head,
            attn_softclamp_value=attn_softclamp_value, max_steps=max_steps, num_discrete_actions=nda,
            multi_token_pred_len=multi_token_pred_len, policy_head_mlp_depth=policy_head_mlp_depth,
            value_head_mlp_depth=value_head_mlp_depth, terminal_mlp_depth=(predict_terminal_mlp_kwargs or {}).get('depth', 1),
            ff_activation=act, ff_expansion_factor=ffk.get('expansion_factor', 4.),
            reward_range=tuple(rk.get('reward_range', (-20., 20.))), reward_num_bins=rk.get('num_bins', 255),
            value_range=tuple(vk.get('reward_range', (-20., 20.))), value_num_bins=vk.get('num_bins', 255),
            predict_terminals=predict_terminals, num_tasks=num_tasks, num_agents=num_agents)
        self.dim, self.depth, self.max_steps = dim, depth, max_steps
        self.predict_terminals = predict_terminals
        self.precision, self.time_attn_variant = precision, time_attn_variant
        self.gae_discount_factor, self.gae_lambda = gae_discount_factor, gae_lambda
        self.ppo_eps_clip, self.policy_entropy_weight = ppo_eps_clip, policy_entropy_weight
        self.use_delight_gating, self.delight_temperature = use_delight_gating, delight_temperature
        self.normalize_advantages = normalize_advantages
        self.pmpo_pos_to_neg_weight, self.pmpo_reverse_kl = pmpo_pos_to_neg_weight, pmpo_reverse_kl     # reference :5227-5231
        self.pmpo_kl_div_loss_weight = pmpo_kl_div_loss_weight
        self.keep_reward_ema_stats, self.reward_ema_decay = keep_reward_ema_stats, reward_ema_decay      # reference :5238-5246
        # knobs of the world-model training branch of forward() (reference :4698-4723, 4898, 5258-5263); forward-only here
        self.loss_weight_fn = kwargs.get('loss_weight_fn', lambda times, slope=0.9, intercept=0.1: slope * times + intercept)     # ramp_weight (:897-899)
        self.num_step_sizes_log2 = int(math.log2(max_steps))
        self.prob_shortcut_train = default(kwargs.get('prob_shortcut_train'), 1. - self.num_step_sizes_log2 ** -1.)
        self.latent_flow_loss_weight, self.shortcut_loss_weight = kwargs.get('latent_flow_loss_weight', 1.), kwargs.get('shortcut_loss_weight', 1.)
        self._loss_weights = {k: kwargs.get(k, 1.) for k in ('reward_loss_weight', 'terminal_loss_weight', 'discrete_action_loss_weight', 'continuous_action_loss_weight')}
        self._reward_quantile_filter = tuple(float(q) for q in reward_quantile_filter)
        self.latent_shape = (num_latent_tokens, dim_latent)
        self._build_parameters()
        self.video_tokenizer = video_tokenizer.eval() if exists(video_tokenizer) else None      # submodule: state_dict keys video_tokenizer.* (4794)
        self._ctx = None
        self._ctx_key = None
        self._kv_epoch = 0
        self._packed = None
        self._packed_version = None
        self._bufs = {}

    # ------------------------------------------------------------------ parameters (reference state_dict layout)

    def _reg(self, path, tensor, buffer=False, persistent=True):
        parts = path.split('.')
        mod = self
        for p in parts[:-1]:
            if p not in mod._modules:
                mod.add_module(p, _Node())
            mod = mod._modules[p]
        if buffer:
            mod.register_buffer(parts[-1], tensor, persistent=persistent)
        else:
            mod.register_parameter(parts[-1], nn.Parameter(tensor))

    def _reg_attention(self, p, dim, dim_kv_in, heads, query_heads, dim_head, norm_context, value_residual):
        self._reg(p + 'norm.weight', torch.ones(dim))
        if norm_context:
            self._reg(p + 'norm_context.weight', torch.ones(dim_kv_in))
        self._reg(p + 'to_q.weight', _linear_w(query_heads * dim_head, dim))
        self._reg(p + 'to_k.weight', _linear_w(heads * dim_head, dim_kv_in))
        self._reg(p + 'to_v.weight', _linear_w(heads * dim_head, dim_kv_in))
        self._reg(p + 'to_out.weight', _linear_w(dim, query_heads * dim_head))
        self._reg(p + 'to_gates.0.weight', _linear_w(query_heads, dim))
        self._reg(p + 'k_heads_rmsnorm.gamma', torch.zeros(heads, dim_head))
        if value_residual:
            self._reg(p + 'to_learned_value_residual_mix.0.weight', _linear_w(heads, dim))
            self._reg(p + 'to_learned_value_residual_mix.0.bias', _linear_b(heads, dim))

    def _reg_ff(self, p, dim, inner):
        self._reg(p + 'norm.weight', torch.ones(dim))
        self._reg(p + 'proj_in.weight', _linear_w(2 * inner, dim))
        self._reg(p + 'proj_in.bias', _linear_b(2 * inner, dim))
        self._reg(p + 'proj_out.weight', _linear_w(dim, inner))
        self._reg(p + 'proj_out.bias', _linear_b(dim, inner))

    def _reg_mlp(self, p, dims):
        n = len(dims) - 1
        for l in range(n):
            self._reg(f'{p}.layers.{l}.0.weight', _linear_w(dims[l + 1], dims[l]))
            self._reg(f'{p}.layers.{l}.0.bias', _linear_b(dims[l + 1], dims[l]))
            if l < n - 1:
                self._reg(f'{p}.layers.{l}.1.weight', torch.ones(dims[l + 1]))
                self._reg(f'{p}.layers.{l}.1.bias', torch.zeros(dims[l + 1]))

    def _build_parameters(self):
        import weakref
        c = self.cfg
        D, Dl, h, hq, d = c.dim, c.dim_latent, c.attn_heads, c.query_heads, c.attn_dim_head
        for which, name in enumerate(('policy_head', 'value_head')):
            node = _HeadNode()
            object.__setattr__(node, '_d4_model', weakref.ref(self))
            object.__setattr__(node, '_d4_which', which)
            self.add_module(name, node)
        self._reg('register_tokens', torch.randn(c.num_register_tokens, D) * 1e-2)
        self._reg('agent_learned_embed', torch.randn(c.num_agents, D) * 1e-2)
        self._reg('action_learned_embed', torch.randn(c.num_agents, D) * 1e-2)
        self._reg('reward_learned_embed', torch.randn(c.num_agents, D) * 1e-2)
        self._reg('latent_genes', torch.randn(0, D) * 1e-2)
        self._reg('ema_returns_mean', torch.zeros(()), buffer=True)          # reference dreamer4.py:5245-5246
        self._reg('ema_returns_var', torch.ones(()), buffer=True)
        self._reg('reward_quantile_filter', torch.tensor(self._reward_quantile_filter), buffer=True, persistent=False)
        for name in ('reward_loss_weight', 'terminal_loss_weight', 'discrete_action_loss_weight', 'continuous_action_loss_weight'):
            self._reg(name, torch.tensor(self._loss_weights[name], dtype=torch.float32), buffer=True)         # reference :5260-5263
        if c.same_len:          # reference dreamer4.py:4819-4834
            self._reg('latents_to_spatial_tokens.weight', _linear_w(D, Dl))
            self._reg('latents_to_spatial_tokens.bias', _linear_b(D, Dl))
        else:
            self._reg('latents_to_spatial_tokens.queries', torch.randn(c.num_spatial_tokens, D) * 1e-2)
            self._reg_attention('latents_to_spatial_tokens.attn.', D, Dl, h, hq, d, True, False)
        self._reg('to_latent_pred.0.weight', torch.ones(D))
        if not c.same_len:
            self._reg('to_latent_pred.1.queries', torch.randn(c.num_latent_tokens, D) * 1e-2)
            self._reg_attention('to_latent_pred.1.attn.', D, D, h, hq, d, True, False)
        self._reg('to_latent_pred.2.weight', _linear_w(Dl, D))
        self._reg('signal_levels_embed.weight', torch.randn(c.max_steps, D // 2))
        self._reg('step_size_embed.weight', torch.randn(int(math.log2(c.max_steps)), D // 2))
        self._reg('task_embed.weight', torch.randn(c.num_tasks, D))
        hid = 4 * D
        self._reg_mlp('policy_head', (D, *((hid,) * (c.policy_head_mlp_depth + 1)), hid))
        A = c.total_actions
        self._reg('action_embedder.discrete_action_unembed', torch.randn(A, c.multi_token_pred_len, hid) * 1e-2)
        self._reg('action_embedder.continuous_action_unembed', torch.randn(0, c.multi_token_pred_len, hid, 2) * 1e-2)
        self._reg('action_embedder.discrete_action_embed.weight', torch.randn(A, D))
        self._reg('action_embedder.continuous_action_embed.weight', torch.randn(0, D))
        for e in range(c.multi_token_pred_len):
            self._reg(f'to_reward_pred.nets.{e}.0.weight', torch.ones(D))
            self._reg(f'to_reward_pred.nets.{e}.1.weight', _linear_w(c.reward_num_bins, D))
        if c.predict_terminals:
            th = 4 * Dl
            self._reg_mlp('to_state_terminal_pred.0', (Dl, *((th,) * (c.terminal_mlp_depth + 1)), 1))
        self._reg_mlp('value_head', (D, *((hid,) * (c.value_head_mlp_depth + 1)), c.value_num_bins))
        inv_freq = 1.0 / (10000. ** (torch.arange(0, d, 2).float() / d))          # Rotary1D, reference dreamer4.py:1604-1612
        self._reg('transformer.time_rotary.inv_freq', inv_freq, buffer=True)
        self._reg('transformer.to_value_residual.0.weight', torch.ones(D))
        self._reg('transformer.to_value_residual.1.weight', _linear_w(h * d, D))
        for i in range(c.depth):
            self._reg_attention(f'transformer.layers.{i}.2.fn.', D, D, h, hq, d, False, True)
            self._reg_ff(f'transformer.layers.{i}.3.fn.', D, c.ff_inner)
        for i in range(c.depth - 1):
            self._reg_attention(f'transformer.attn_pools.{i}.fn.attn.', D, D, c.pool_heads, c.pool_heads, c.pool_dim_head, True, False)
        self._reg_attention('transformer.final_attn_pool.fn.attn.', D, D, c.pool_heads, c.pool_heads, c.pool_dim_head, True, False)
        self._reg_attention('transformer.final_special_cross_attn.fn.', D, D, h, hq, d, True, True)
        self._reg_ff('transformer.final_special_ff.fn.', D, c.ff_inner)

    @property
    def device(self):
        return self.register_tokens.device

    def _named(self, prefixes):
        return [p for n, p in self.named_parameters() if n.startswith(prefixes)]

    def policy_head_parameters(self):      # reference dreamer4.py:5343-5352: the policy MLP + the action UNembedding (1249-1250), not the embedding
        unembeds = ('action_embedder.discrete_action_unembed', 'action_embedder.continuous_action_unembed')
        return self._named(('policy_head.',)) + [p for n, p in self.named_parameters() if n in unembeds]

    def muon_parameters(self):             # reference dreamer4.py:5335-5341, 2918-2925, 1960-1966, 2099-2103: to_v / to_out / proj_in / proj_out weights
        ends = ('.to_v.weight', '.to_out.weight', '.proj_in.weight', '.proj_out.weight')
        return [p for n, p in self.named_parameters() if n.startswith('transformer.') and n.endswith(ends)]

    def value_head_parameters(self):       # reference dreamer4.py:5354-5363
        return self._named(('value_head.',))

    # .save / .load / .init_and_load of the reference's @save_load (dreamer4.py:4660; used at cli.py:329, tests/test_dreamer.py:2243-2247):
    # one torch.save'd dict {model: state_dict, config: pickled (args, kwargs)}

    def save(self, path, overwrite=True):
        import os
        assert overwrite or not os.path.exists(str(path)), f'{path} already exists'
        torch.save(dict(model=self.state_dict(), config=pickle.dumps(self._config)), str(path))

    def load(self, path, strict=True):
        pkg = torch.load(str(path), map_location='cpu', weights_only=False)
        self.load_state_dict(pkg['model'], strict=strict)

    @classmethod
    def init_and_load(cls, path, strict=True):
        pkg = torch.load(str(path), map_location='cpu', weights_only=False)
        args, kwargs = pickle.loads(pkg['config'])
        model = cls(*args, **kwargs)
        model.load_state_dict(pkg['model'], strict=strict)
        return model

    # ------------------------------------------------------------------ engine plumbing

    def _require_cuda(self):
        if self.device.type != 'cuda':
            raise D4Error('dreamer4_b200 runs on CUDA only: move the model to a B200 (`model.cuda()`); there is no CPU fallback')

    def _frozen_version(self):
        return sum(p._version for n, p in self.named_parameters() if not n.startswith(('policy_head.', 'value_head.', 'to_state_terminal_pred.'))
                   and n != 'action_embedder.discrete_action_unembed')

    def _engine(self, batch, max_time, agent_index=0, grow=False):
        """(Re)creates the native context for this (batch, max_time) capacity and binds weights.  `grow` (prompted /
        resumed rollouts, whose time_steps creeps up by one per env step): an existing context whose KV capacity already
        covers max_time is kept, a new one is sized to the next multiple of 64 frames."""
        self._require_cuda()
        lib = _lib.load()
        c = self.cfg
        dev = self.device
        # an existing context of the same batch whose KV capacity already covers max_time is kept (no re-allocation, captured frame
        # graphs stay valid: env-style stepping alternates 1-frame resets with growing prompted calls); otherwise a new one is sized
        # exactly (`grow` False: the DreamTrainer path keeps its exact-size KV buffer) or to the next multiple of 64 frames
        k = self._ctx_key
        if self._ctx is not None and k[0] == batch and k[1] >= max_time and k[2:] == (agent_index, self.precision, self.time_attn_variant, dev.index):
            max_time = k[1]
        elif grow:
            max_time = (max_time + 63) // 64 * 64
        key = (batch, max_time, agent_index, self.precision, self.time_attn_variant, dev.index)
        if self._ctx is not None and self._ctx_key != key:
            self._release()
        if self._ctx is None:
            cc = _lib.d4_config()
            cc.dim, cc.dim_latent, cc.num_latent_tokens = c.dim, c.dim_latent, c.num_latent_tokens
            cc.num_spatial_tokens, cc.num_register_tokens = c.num_spatial_tokens, c.num_register_tokens
            cc.depth, cc.time_block_every = c.depth, c.time_block_every
            cc.heads, cc.query_heads, cc.dim_head = c.attn_heads, c.query_heads, c.attn_dim_head
            cc.pool_heads, cc.pool_dim_head = c.pool_heads, c.pool_dim_head
            cc.ff_inner, cc.ff_inner_pad, cc.ff_act = c.ff_inner, c.ff_inner_pad, 1 if c.ff_activation == 'gelu' else 0
            cc.max_steps = c.max_steps
            cc.num_action_types = len(c.num_discrete_actions)
            for i, n in enumerate(c.num_discrete_actions):
                cc.action_sizes[i] = n
            cc.policy_layers, cc.policy_hidden = c.policy_head_mlp_depth + 2, 4 * c.dim
            cc.value_layers, cc.value_hidden = c.value_head_mlp_depth + 2, 4 * c.dim
            cc.terminal_layers, cc.terminal_hidden = c.terminal_mlp_depth + 2, 4 * c.dim_latent
            cc.predict_terminals = int(c.predict_terminals)
            cc.reward_bins, cc.value_bins, cc.num_tasks = c.reward_num_bins, c.value_num_bins, c.num_tasks
            cc.softclamp = c.attn_softclamp_value
            cc.max_batch, cc.max_time = batch, max_time
            cc.precision, cc.time_attn_variant = _lib.PREC[self.precision], self.time_attn_variant
            ctx = C.c_void_p()
            check(lib.d4_ctx_create(C.byref(cc), C.byref(ctx)))
            self._ctx, self._ctx_key = ctx, key
            ws_bytes, kv_bytes = lib.d4_workspace_bytes(ctx), lib.d4_kv_bytes(ctx)
            with torch.cuda.device(dev):
                ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=dev)
                ws = ws[(-ws.data_ptr()) % 256:][:ws_bytes]      # 256-byte aligned (a no-op offset for CUDA allocations)
                y = max(c.num_time_layers, 1)
                kv = torch.zeros(y, 2, batch * c.tokens_per_frame, c.attn_heads, max_time, c.attn_dim_head, device=dev)
                check(lib.d4_set_buffers(ctx, ptr(ws), ws_bytes, ptr(kv), kv.numel() * 4))
            self._bufs = dict(ws=ws, kv=kv)
            self._packed_version = None
        ver = (self._frozen_version(), agent_index)
        if self._packed_version != ver:
            sd = {k: v for k, v in self.state_dict().items()}
            split = self.precision in ('tf32x3', 'f16x3')       # f16x3: d4_learn and odd shapes stay on 3xTF32
            self._packed = pack(sd, c, dev, agent_index=agent_index, split=split, split_f16=self.precision == 'f16x3')
            for name, scale in self._packed.pop('h16scales', {}).items():
                check(lib.d4_set_weight_scale(self._ctx, name.encode(), scale))
            for name, t in self._packed.items():
                check(lib.d4_set_weight(self._ctx, name.encode(), ptr(t), t.numel()))
            params = dict(self.named_parameters())
            heads = [('policy', 'policy_head', c.policy_head_mlp_depth + 2), ('value', 'value_head', c.value_head_mlp_depth + 2)] if c.has_actions else []
            if c.predict_terminals:
                heads.append(('terminal', 'to_state_terminal_pred.0', c.terminal_mlp_depth + 2))
            self._head_splits = []
            for short, prefix, nl in heads:
                for pname, key_ in mlp_param_names(prefix, nl):
                    t = params[key_]
                    assert t.is_contiguous()
                    check(lib.d4_set_weight(self._ctx, f'{short}.{pname}'.encode(), ptr(t), t.numel()))
                    if split and pname.endswith('.w') and short != 'terminal':
                        # trained head weights change every optimizer step: their tf32 hi/lo split lives in persistent
                        # buffers that _refresh_head_splits() rewrites in place when the parameter version moves
                        hi, lo = torch.empty_like(t), torch.empty_like(t)
                        thi, tlo = torch.empty_like(t.t().contiguous()), torch.empty_like(t.t().contiguous())     # of W^T: learn backward
                        h16 = None
                        if self.precision == 'f16x3':
                            # the rollout's head MLPs also run on the fp16 split GEMM: fp16 words of q w with q a power of two fixed here
                            # (rms(q w) ~ 1); _refresh_head_splits() re-splits with the same q and re-packs if the weights outgrow it
                            rms = float(t.detach().float().pow(2).mean().sqrt().clamp_min(1e-30))
                            q = 2.0 ** round(math.log2(1.0 / rms))
                            h16 = dict(hi=torch.empty_like(t, dtype=torch.float16), lo=torch.empty_like(t, dtype=torch.float16), q=q, name=f'{short}.{pname}')
                            for suffix, buf in (('.h16hi', h16['hi']), ('.h16lo', h16['lo'])):
                                check(lib.d4_set_weight(self._ctx, f'{short}.{pname}{suffix}'.encode(), ptr(buf), buf.numel()))
                            check(lib.d4_set_weight_scale(self._ctx, h16['name'].encode(), 1.0 / q))
                        self._head_splits.append((t, hi, lo, thi, tlo, h16))
                        for suffix, buf in (('.hi', hi), ('.lo', lo)):
                            check(lib.d4_set_weight(self._ctx, f'{short}.{pname}{suffix}'.encode(), ptr(buf), buf.numel()))
                        for suffix, buf in (('.hi', thi), ('.lo', tlo)):
                            check(lib.d4_set_weight(self._ctx, f'{short}.{pname[:-2]}.wt{suffix}'.encode(), ptr(buf), buf.numel()))
            self._head_split_version = None
            if c.has_actions:
                un = params['action_embedder.discrete_action_unembed']          # (A, mtp, 4D): head 0 rows, row stride mtp*4D
                check(lib.d4_set_weight(self._ctx, b'unembed', ptr(un), un.stride(0)))
            check(lib.d4_bind(self._ctx))
            self._packed_version = ver
        self._refresh_head_splits()
        return lib, self._ctx

    @torch.no_grad()
    def _refresh_head_splits(self):
        splits = getattr(self, '_head_splits', None)
        if not splits:
            return
        ver = sum(s[0]._version for s in splits)
        if ver == self._head_split_version:
            return
        rebind = False
        for t, hi, lo, thi, tlo, h16 in splits:
            h, l = tf32_split(t.detach())
            hi.copy_(h)
            lo.copy_(l)
            thi.copy_(hi.t())
            tlo.copy_(lo.t())
            if h16 is not None:
                wq = t.detach().float() * h16['q']
                if float(wq.abs().max()) > 16384.:          # the weights outgrew the power of two chosen at pack time: choose again, re-bind
                    rms = float(t.detach().float().pow(2).mean().sqrt().clamp_min(1e-30))
                    h16['q'] = 2.0 ** round(math.log2(1.0 / rms))
                    wq = t.detach().float() * h16['q']
                    check(_lib.load().d4_set_weight_scale(self._ctx, h16['name'].encode(), 1.0 / h16['q']))
                    rebind = True
                h16['hi'].copy_(wq)
                h16['lo'].copy_(wq - h16['hi'].float())
        if rebind:
            check(_lib.load().d4_bind(self._ctx))
        self._head_split_version = ver

    def _release(self):
        if self._ctx is not None:
            _lib.load().d4_ctx_destroy(self._ctx)
        self._ctx, self._ctx_key, self._bufs, self._packed, self._packed_version = None, None, {}, None, None
        self._head_splits, self._head_split_version = [], None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    _NATIVE_STATE = dict(_ctx=None, _ctx_key=None, _bufs={}, _packed=None, _packed_version=None, _head_splits=[], _head_split_version=None)

    def __deepcopy__(self, memo):
        """A copy owns no native context (the handles borrow this instance's parameter storage): it builds its own on first use."""
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        new.__dict__ = {k: (copy.deepcopy(self._NATIVE_STATE[k]) if k in self._NATIVE_STATE else copy.deepcopy(v, memo)) for k, v in self.__dict__.items()}
        import weakref
        for name in ('policy_head', 'value_head'):          # the callable heads point back at their owner
            object.__setattr__(new._modules[name], '_d4_model', weakref.ref(new))
        return new

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._release()           # parameter storage moved: borrowed pointers are stale
        return out

    @torch.no_grad()
    def _head_forward(self, which, x):
        """x (..., dim_in) -> (..., dim_out) through head `which` (0 policy, 1 value, 2 terminal) on the native MLP kernels."""
        self._require_cuda()
        c = self.cfg
        assert c.has_actions or which == 2, 'the model has no policy / value head (no actions)'
        dim_in = c.dim_latent if which == 2 else c.dim
        dim_out = (4 * c.dim, c.value_num_bins, 1)[which]
        assert x.shape[-1] == dim_in, f'head input has {x.shape[-1]} features, expected {dim_in}'
        rows = x.reshape(-1, dim_in).to(device=self.device, dtype=torch.float32).contiguous()
        lib, ctx = self._engine(*(self._ctx_key[:2] if self._ctx_key else (min(max(rows.shape[0], 1), 4096), 1)))
        cap = self._ctx_key[0]
        out = torch.empty(rows.shape[0], dim_out, device=self.device, dtype=torch.float32)
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        for r0 in range(0, rows.shape[0], cap):
            m = min(cap, rows.shape[0] - r0)
            check(lib.d4_head_forward(ctx, which, C.c_void_p(rows[r0:].data_ptr()), m, C.c_void_p(out[r0:].data_ptr()), stream))
        return out.reshape(*x.shape[:-1], dim_out)

    # ------------------------------------------------------------------ forward (inference branch)

    @torch.no_grad()
    def forward(self, *, video=None, latents=None, lens=None, signal_levels=None, step_sizes=None, step_sizes_log2=None, tasks=None,
                rewards=None, terminals=None, discrete_actions=None, continuous_actions=None, shift_action_tokens=True, proprio=None,
                time_cache=None, return_pred_only=False, latent_is_noised=False, return_all_losses=False, return_intermediates=False,
                latent_has_view_dim=False, agent_index=0, seed=None, **unsupported):
        """The inference branch of the reference's forward (dreamer4.py:6792-7295): `signal_levels` and a step size given, the
        prediction (and with `return_intermediates` the agent embeddings and the time cache) returned - what `generate`,
        `interact_with_env` and the reference's parallel-vs-sequential tests (tests/test_dreamer.py:1206-1296) call.  Runs frame by
        frame over the in-place time-KV cache: time attention is causal, so a T-frame call equals the reference's uncached
        multi-frame forward, and a call with `time_cache` continues at its cache position.

        latents (b, t, n, d) [or (b, t, 1, n, d) with latent_has_view_dim]; signal_levels int | (b,) | (b, t); step_sizes /
        step_sizes_log2 int | (b,); discrete_actions (b, t, na) | (b, t-1, na) | None, shifted as the reference does (:7105-7126).
        Returns Predictions(flow (b, t, 1, n, d)) or (Predictions, (Embeds(agent (b, t, 1, D)), DynamicsIntermediates)).

        The training branch (no signal_levels: flow / shortcut / reward / action losses, :7297-7743) is a "next" row."""
        unsupported_noise, unsupported_schedule = unsupported.pop('noise', None), unsupported.pop('train_schedule', None)      # injected draws (tests)
        add_ar_action_loss = unsupported.pop('add_autoregressive_action_loss', True)
        unsupported.pop('update_loss_ema', None)
        for name, v in dict(proprio=proprio, continuous_actions=continuous_actions, **unsupported).items():
            if exists(v) and v is not False:
                raise NotImplementedError(f'forward({name}=...) is outside the B200 hot path (SURVEY.md section 8)')
        assert exists(video) ^ exists(latents)
        if exists(video):
            assert exists(self.video_tokenizer), 'video_tokenizer must be passed in if training from raw video on dynamics model'
            latents = self.video_tokenizer.tokenize(video)
        is_inference = exists(signal_levels)
        assert not (exists(signal_levels) ^ (exists(step_sizes) or exists(step_sizes_log2))), 'signal_levels and a step size go together'
        if not is_inference or not (return_pred_only or latent_is_noised):
            return self._training_forward(latents, lens=lens, signal_levels=signal_levels, step_sizes=step_sizes, step_sizes_log2=step_sizes_log2,
                                          tasks=tasks, rewards=rewards, terminals=terminals, discrete_actions=discrete_actions,
                                          shift_action_tokens=shift_action_tokens, return_all_losses=return_all_losses, latent_has_view_dim=latent_has_view_dim,
                                          agent_index=agent_index, seed=seed, noise=unsupported_noise, schedule=unsupported_schedule,
                                          add_autoregressive_action_loss=add_ar_action_loss)
        c = self.cfg
        dev = self.device
        if latents.ndim == 5:
            assert latent_has_view_dim and latents.shape[2] == 1, 'multi-view latents are outside the B200 hot path'
            latents = latents[:, :, 0]
        assert tuple(latents.shape[-2:]) == self.latent_shape, f'latents must have shape {self.latent_shape}, got {tuple(latents.shape[-2:])}'
        B, T = latents.shape[:2]
        N, Dl, D, na = c.num_latent_tokens, c.dim_latent, c.dim, len(c.num_discrete_actions)
        f32 = dict(device=dev, dtype=torch.float32)
        latents = latents.to(**f32)

        def per_batch(v, name):
            v = torch.as_tensor(v, device=dev)
            assert v.ndim <= 1, f'{name} must be a scalar or (b,)'
            return v.expand(B) if v.ndim == 0 else v
        sig = torch.as_tensor(signal_levels, device=dev)
        sig = sig.expand(B) if sig.ndim == 0 else sig
        sig = sig[:, None].expand(B, T) if sig.ndim == 1 else sig
        assert tuple(sig.shape) == (B, T), f'signal_levels {tuple(sig.shape)}'
        assert not (exists(step_sizes) and exists(step_sizes_log2))
        if exists(step_sizes):                                   # reference :6936-6942
            ss = per_batch(step_sizes, 'step_sizes').float()
            log2 = torch.log2(ss)
            step_log2 = log2.long()
            assert bool((step_log2 == log2).all()), '`step_sizes` must be powers of 2'
        else:
            step_log2 = per_batch(step_sizes_log2, 'step_sizes_log2').long()
        sig = sig.long().contiguous()
        step_log2 = step_log2.contiguous()
        assert int(sig.min()) >= 0 and int(sig.max()) < self.max_steps and int(step_log2.min()) >= 0 and \
            int(step_log2.max()) < int(math.log2(self.max_steps)), 'signal level / step size outside the embedding tables'
        if not latent_is_noised:                                 # reference :6990-7001: noise.lerp(latents, times), times = signal / max_steps
            gen = torch.Generator(device=dev).manual_seed(seed) if exists(seed) else None
            times = (sig.float() / self.max_steps)[..., None, None]
            latents = torch.randn(latents.shape, generator=gen, **f32).lerp(latents, times)
        latents = latents.contiguous()

        # action tokens (reference :7088-7126): frame i is conditioned on `prev[i]` (None: the zero token)
        prev = [None] * T
        if exists(discrete_actions):
            assert c.has_actions
            acts = discrete_actions if discrete_actions.ndim == 3 else discrete_actions[..., None]
            acts = acts.to(dev, torch.long).contiguous()
            alen = acts.shape[1]
            sequential = exists(time_cache) and T == 1 and alen == 1
            if alen == T and shift_action_tokens and not sequential:
                prev = [None] + [acts[:, i] for i in range(T - 1)]
            elif alen == T - 1:
                prev = [None] + [acts[:, i] for i in range(T - 1)]
            else:
                assert alen == T, f'discrete_actions cover {alen} steps for {T} frames'
                prev = [acts[:, i] for i in range(T)]
        if isinstance(tasks, int):
            tasks = torch.full((B,), tasks, device=dev, dtype=torch.long)
        if exists(tasks):
            tasks = tasks.to(dev, torch.long).contiguous()

        resumed_kv, t0 = None, 0
        if exists(time_cache) and exists(time_cache.main):
            resumed_kv, t0 = time_cache.main.next_kv_cache, time_cache.main.token_count
        lib, ctx, kv = self._adopt_time_cache(resumed_kv, t0, B, t0 + T, agent_index, grow=True)
        self._kv_epoch += 1
        flow = torch.empty(B, T, N, Dl, **f32)
        agent = torch.empty(B, T, D, **f32)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        for i in range(T):
            frame, s_i = latents[:, i].contiguous(), sig[:, i].contiguous()
            pa = prev[i].contiguous() if exists(prev[i]) else None
            pred_i, agent_i = torch.empty(B, N, Dl, **f32), torch.empty(B, D, **f32)
            check(lib.d4_pass_ex(ctx, B, ptr(frame), ptr(s_i), ptr(step_log2), ptr(pa), na, ptr(tasks), t0 + i, 1, ptr(pred_i), ptr(agent_i), stream))
            flow[:, i], agent[:, i] = pred_i, agent_i
        pred = Predictions(flow[:, :, None], None, None)
        if not return_intermediates:
            return pred
        L = c.num_time_layers
        next_kv = None
        if L > 0:
            next_kv = kv[:L, :, :, :, :t0 + T]
            next_kv._d4_epoch = self._kv_epoch
        inter = DynamicsIntermediates(main=TransformerIntermediates(next_kv_cache=next_kv, token_count=t0 + T))
        return pred, (Embeds(agent=agent[:, :, None]), inter)

    # ------------------------------------------------------------------ forward (training branch, forward only)

    def _frames_pass(self, latents, sig, step_log2, prev, tasks, agent_index):
        """One cache-less multi-frame forward: frames 0..T-1 of `latents` (b, t, n, d) at signal levels `sig` (b, t) through d4_pass_ex,
        each committing its keys / values at cache position i (causal time attention = the reference's uncached parallel forward).
        Returns (prediction (b, t, n, d), agent embeddings (b, t, D))."""
        c = self.cfg
        B, T = latents.shape[:2]
        dev = self.device
        f32 = dict(device=dev, dtype=torch.float32)
        lib, ctx, _ = self._adopt_time_cache(None, 0, B, T, agent_index, grow=True)
        self._kv_epoch += 1
        flow, agent = torch.empty(B, T, c.num_latent_tokens, c.dim_latent, **f32), torch.empty(B, T, c.dim, **f32)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        na = len(c.num_discrete_actions)
        for i in range(T):
            frame, s_i = latents[:, i].contiguous(), sig[:, i].contiguous()
            pa = prev[i].contiguous() if exists(prev[i]) else None
            pred_i, agent_i = torch.empty(B, c.num_latent_tokens, c.dim_latent, **f32), torch.empty(B, c.dim, **f32)
            check(lib.d4_pass_ex(ctx, B, ptr(frame), ptr(s_i), ptr(step_log2), ptr(pa), na, ptr(tasks), i, 1, ptr(pred_i), ptr(agent_i), stream))
            flow[:, i], agent[:, i] = pred_i, agent_i
        return flow, agent

    @torch.no_grad()
    def _training_forward(self, latents, *, lens, signal_levels, step_sizes, step_sizes_log2, tasks, rewards, terminals, discrete_actions,
                          shift_action_tokens, return_all_losses, latent_has_view_dim, agent_index, seed, noise, schedule,
                          add_autoregressive_action_loss):
        """World-model training losses, FORWARD ONLY (reference dreamer4.py:6963-6997, 7297-7743): flow loss with the
        ramp weighting, shortcut (consistency) loss from two extra half-step passes, multi-token reward / discrete-action prediction
        losses, terminal loss, and their weighted total.  The transformer passes - all of the arithmetic that matters - and the head
        MLPs run on the native kernels (d4_pass_ex, d4_head_forward, d4_linear); the loss assembly on their (b, t, ...) outputs is a
        handful of torch reductions.  No gradients: the returned scalars carry no autograd graph (the backward of the transformer is
        outside this package).  `noise` / `train_schedule` (not in the reference) inject the random draws of the branch."""
        from .experience import WorldModelLosses
        c = self.cfg
        dev = self.device
        f32 = dict(device=dev, dtype=torch.float32)
        if latents.ndim == 5:
            assert latent_has_view_dim and latents.shape[2] == 1, 'multi-view latents are outside the B200 hot path'
            latents = latents[:, :, 0]
        assert tuple(latents.shape[-2:]) == self.latent_shape, f'latents must have shape {self.latent_shape}, got {tuple(latents.shape[-2:])}'
        latents = latents.to(**f32).contiguous()
        B, T = latents.shape[:2]
        N, Dl, D, na, K = c.num_latent_tokens, c.dim_latent, c.dim, len(c.num_discrete_actions), c.multi_token_pred_len
        if exists(rewards):                                                           # reference :6897-6901
            rewards = rewards.to(**f32)
            if rewards.shape[1] == T - 1:
                rewards = torch.nn.functional.pad(rewards, (1, 0), value=0.)
            assert rewards.shape[1] == T, f'during training, rewards must perfectly align with video length {T}, got {rewards.shape[1]}'
        if exists(terminals):
            assert terminals.ndim == 2, 'terminals (b, t) or (b, t-1) per-frame flags (a (b,) tensor trips the reference at :6904 too)'
            terminals = terminals.to(dev)
            if terminals.shape[1] == T - 1:
                terminals = torch.nn.functional.pad(terminals, (1, 0), value=False)
        if isinstance(tasks, int):
            tasks = torch.full((B,), tasks, device=dev, dtype=torch.long)
        if exists(tasks):
            tasks = tasks.to(dev, torch.long).contiguous()

        # ---- signal levels and step sizes: given, injected, or drawn as the reference draws them (:6963-6979)
        is_inference = exists(signal_levels)
        gen = torch.Generator(device=dev).manual_seed(seed) if exists(seed) else None
        shortcut_train = False
        if is_inference:
            sig = torch.as_tensor(signal_levels, device=dev)
            sig = sig.expand(B) if sig.ndim == 0 else sig
            sig = (sig[:, None].expand(B, T) if sig.ndim == 1 else sig).long()
            if exists(step_sizes):
                step_log2 = torch.log2(torch.as_tensor(step_sizes, device=dev).float()).long()
            else:
                step_log2 = torch.as_tensor(step_sizes_log2, device=dev).long()
            step_log2 = step_log2.expand(B) if step_log2.ndim == 0 else step_log2
        elif exists(schedule):
            shortcut_train = bool(schedule['shortcut_train'])
            step_log2, sig = schedule['step_sizes_log2'].to(dev).long(), schedule['signal_levels'].to(dev).long()
        else:
            shortcut_train = bool(torch.rand(1, device=dev, generator=gen).item() < self.prob_shortcut_train)
            if shortcut_train:
                step_log2 = torch.randint(1, self.num_step_sizes_log2, (B,), device=dev, generator=gen)
                nss = 2 ** step_log2
                sig = torch.randint(0, self.max_steps, (B, T), device=dev, generator=gen) // nss[:, None] * nss[:, None]
            else:
                step_log2 = torch.zeros(B, device=dev, dtype=torch.long)
                sig = torch.randint(0, self.max_steps, (B, T), device=dev, generator=gen)
        sig, step_log2 = sig.contiguous(), step_log2.contiguous()
        times = sig.float() / self.max_steps                                           # (b, t)
        t4 = times[..., None, None]
        noise = noise.to(**f32) if exists(noise) else torch.randn(latents.shape, generator=gen, **f32)
        noised = noise.lerp(latents, t4)                                               # :6995-7001

        # ---- action tokens (:7088-7126) and the three passes
        prev = [None] * T
        acts = None
        if exists(discrete_actions):
            assert c.has_actions
            acts = (discrete_actions if discrete_actions.ndim == 3 else discrete_actions[..., None]).to(dev, torch.long).contiguous()
            alen = acts.shape[1]
            if (alen == T and shift_action_tokens) or alen == T - 1:
                prev = [None] + [acts[:, i] for i in range(T - 1)]
            else:
                assert alen == T, f'discrete_actions cover {alen} steps for {T} frames'
                prev = [acts[:, i] for i in range(T)]
        pred, agent = self._frames_pass(noised, sig, step_log2, prev, tasks, agent_index)
        flow_losses = (pred - latents).square() * self.loss_weight_fn(times)[..., None, None]      # x-space: the target is the data (:7343-7350, 7413-7417)

        shortcut_losses = None
        if (not is_inference) and shortcut_train:                                       # :7354-7406
            half_log2 = step_log2 - 1
            half = 2 ** half_log2
            first_pred, _ = self._frames_pass(noised, sig, half_log2, prev, tasks, agent_index)
            first_flow = (first_pred - noised) / (1. - t4)
            denoised = noised + first_flow * (half.float() / self.max_steps)[:, None, None, None]
            sig2 = sig + half[:, None]
            second_pred, _ = self._frames_pass(denoised, sig2.contiguous(), half_log2, prev, tasks, agent_index)
            second_flow = (second_pred - denoised) / (1. - (sig2.float() / self.max_steps)[..., None, None])
            target = (first_flow + second_flow) / 2
            shortcut_pred = (pred - noised) / (1. - t4)
            shortcut_losses = (shortcut_pred - target).square() * (1. - t4) ** 2

        zero = torch.zeros((), **f32)
        loss_mask = None
        if exists(lens):                                                                # :7421-7433
            loss_mask = torch.arange(T, device=dev)[None, :] < lens.to(dev)[:, None]
            flow_loss = flow_losses[loss_mask].mean()
            shortcut_loss = shortcut_losses[loss_mask].mean() if exists(shortcut_losses) else zero
        else:
            flow_loss = flow_losses.mean()
            shortcut_loss = shortcut_losses.mean() if exists(shortcut_losses) else zero
        mask_wo_last = loss_mask[:, :-1] if exists(loss_mask) else None

        def mtp_targets(t, steps):                                                      # create_multi_token_prediction_targets (:530-552)
            L = t.shape[1]
            idx = torch.arange(L, device=dev)[:, None] + torch.arange(steps, device=dev)[None, :]
            ok = idx < L
            return t[:, idx.masked_fill(~ok, 0)], ok[None].expand(t.shape[0], -1, -1)

        # ---- reward loss (:7447-7463): every prediction head on the agent token of the PREVIOUS frame against HL-Gauss(reward)
        reward_loss = zero
        if exists(rewards) and T > 1:
            support, _ = hl_gauss_tables(*c.reward_range, c.reward_num_bins, dev)
            sigma_sqrt2 = math.sqrt(2.) * c.hl_gauss_sigma_to_bin_ratio * (c.reward_range[1] - c.reward_range[0]) / c.reward_num_bins
            cdf = torch.special.erf((support - rewards.clamp(*c.reward_range)[..., None]) / sigma_sqrt2)
            two_hot = (cdf[..., 1:] - cdf[..., :-1]) / (cdf[..., -1] - cdf[..., 0]).clamp(min=c.hl_gauss_eps)[..., None]      # (b, t, bins)
            rows = agent[:, :-1].reshape(-1, D).contiguous()
            rstd = torch.rsqrt(rows.square().mean(dim=-1) + torch.finfo(torch.float32).eps).contiguous()
            sd = dict(self.named_parameters())
            logp = []
            for e in range(K):                                                           # Ensemble of RMSNorm -> Linear (no bias), :5067-5075
                w = (sd[f'to_reward_pred.nets.{e}.1.weight'] * sd[f'to_reward_pred.nets.{e}.0.weight'][None, :]).contiguous()
                out = torch.empty(rows.shape[0], c.reward_num_bins, **f32)
                check(_lib.load().d4_linear(0, rows.shape[0], c.reward_num_bins, D, ptr(rows), D, ptr(w), D, None, None, ptr(rstd), None, 0, 0, ptr(out),
                                            c.reward_num_bins, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
                logp.append(out.log_softmax(dim=-1).reshape(B, T - 1, -1))
            logp = torch.stack(logp, dim=2)                                              # (b, t-1, mtp, bins)
            tgt, ok = mtp_targets(two_hot[:, 1:], K)                                     # (b, t-1, mtp, bins), (b, t-1, mtp)
            reward_losses = (-(tgt * logp).sum(dim=-1)).masked_fill(~ok, 0.)
            reward_loss = reward_losses[mask_wo_last].mean(dim=0) if exists(loss_mask) else reward_losses.mean(dim=(0, 1))

        # ---- terminal loss (:7467-7490)
        terminal_loss = zero
        if exists(terminals) and self.predict_terminals and T > 1:
            pooled = latents[:, 1:].mean(dim=2)                                          # (b, t-1, d)
            logit = self._head_forward(2, pooled)[..., 0]
            eps_t = 1. - self.gae_discount_factor
            tseq = terminals[:, 1:].float().clamp(min=eps_t, max=1. - eps_t)
            tl = torch.nn.functional.binary_cross_entropy_with_logits(logit, tseq, reduction='none')
            terminal_loss = tl[mask_wo_last].mean() if exists(loss_mask) else tl.mean()

        # ---- autoregressive discrete-action loss (:7517-7590)
        action_loss = zero
        if exists(acts) and add_autoregressive_action_loss and T > 1 and c.num_agents == 1 and float(self.discrete_action_loss_weight.sum()) > 0:
            padded = torch.nn.functional.pad(acts, (0, 0, 1, 0), value=-1) if shift_action_tokens else acts
            plen = padded.shape[1]
            num_targets = plen - 1 if shift_action_tokens else plen
            pe = self._head_forward(0, agent[:, :num_targets])                           # (b, nt, 4D)
            un = dict(self.named_parameters())['action_embedder.discrete_action_unembed']      # (A, mtp, 4D)
            lib = _lib.load()
            rows = pe.reshape(-1, pe.shape[-1]).contiguous()
            tgt, ok = mtp_targets(padded, K)                                             # (b, plen, mtp, na)
            if shift_action_tokens:
                tgt, ok = tgt[:, 1:], ok[:, 1:]
            lps = []
            for e in range(K):
                w = un[:, e].contiguous()
                logits = torch.empty(rows.shape[0], c.total_actions, **f32)
                check(lib.d4_linear(0, rows.shape[0], c.total_actions, rows.shape[1], ptr(rows), rows.shape[1], ptr(w), rows.shape[1], None, None, None, None,
                                    0, 0, ptr(logits), c.total_actions, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
                logits = logits.reshape(B, num_targets, -1)
                per_type, off = [], 0
                for ti, n in enumerate(c.num_discrete_actions):
                    lp = logits[..., off:off + n].log_softmax(dim=-1)
                    per_type.append(lp.gather(-1, tgt[:, :, e, ti].clamp(min=0)[..., None])[..., 0])
                    off += n
                lps.append(torch.stack(per_type, dim=-1))                                # (b, nt, na)
            lps = torch.stack(lps, dim=0).masked_fill(~ok.permute(2, 0, 1)[..., None], 0.)      # (mtp, b, nt, na)
            if exists(loss_mask):
                am = mask_wo_last if plen == T - 1 else loss_mask
                action_loss = (-lps).permute(1, 2, 3, 0)[am].mean(dim=(0, 1))
            else:
                action_loss = (-lps).mean(dim=(1, 2, 3))

        total = (flow_loss * self.latent_flow_loss_weight + shortcut_loss * self.shortcut_loss_weight + (reward_loss * self.reward_loss_weight).sum() +
                 terminal_loss * self.terminal_loss_weight + (action_loss * self.discrete_action_loss_weight).sum())
        if not return_all_losses:
            return total
        return total, WorldModelLosses(flow_loss, shortcut_loss, reward_loss, terminal_loss, action_loss, *([zero] * 10))

    # ------------------------------------------------------------------ generate

    def _adopt_time_cache(self, resumed_kv, P, B, T, agent_index, grow):
        """Engine context with KV capacity >= T whose first P frames hold `resumed_kv` (None: nothing to adopt): a view of the
        live in-place buffer is taken as is - if no other rollout has written the buffer since it was handed out - anything
        else is copied in.  Returns (lib, ctx, kv buffer)."""
        c = self.cfg
        lib, ctx = self._engine(B, T, agent_index, grow=grow)
        kv = self._bufs['kv']
        L, BS = c.num_time_layers, B * c.tokens_per_frame
        if exists(resumed_kv) and P > 0 and L > 0:
            dst = kv[:L, :, :BS, :, :P]
            assert tuple(resumed_kv.shape) == tuple(dst.shape), f'time_cache kv {tuple(resumed_kv.shape)} != {tuple(dst.shape)}'
            if resumed_kv.untyped_storage().data_ptr() == kv.untyped_storage().data_ptr():
                # a view of the live in-place cache: only valid if no other rollout has written the buffer since
                if getattr(resumed_kv, '_d4_epoch', None) != self._kv_epoch or resumed_kv.data_ptr() != dst.data_ptr() or resumed_kv.stride() != dst.stride():
                    raise ValueError('stale time_cache: it is a view of the in-place KV buffer, which a later generate() has '
                                     'overwritten; clone next_kv_cache to keep a cache across rollouts')
            else:
                dst.copy_(resumed_kv.to(device=self.device, dtype=torch.float32))
        return lib, ctx, kv

    @torch.no_grad()
    def generate(self, time_steps, num_steps=4, batch_size=1, agent_index=0, tasks=None, latent_gene_ids=None, image_height=None,
                 image_width=None, return_decoded_video=None, context_signal_noise=0.1, time_cache=None, use_time_cache=True,
                 return_rewards_per_frame=False, return_terminals=False, return_agent_actions=False,
                 return_log_probs_and_values=False, return_for_policy_optimization=False, return_time_cache=False,
                 store_agent_embed=True, store_old_action_unembeds=True, prompt=None, prompt_latents=None, prompt_proprio=None,
                 prompt_discrete_actions=None, prompt_continuous_actions=None, prompt_rewards=None, aug_id=False,
                 discrete_temperature=1., continuous_temperature=1., noise=None):
        """Imagination rollout (reference dreamer4.py:6307-6774).  `noise` (not in the reference) optionally injects the
        per-frame random draws — dict(latent=(T,B,N,Dl) normal, action_uniform=(T,B,A_total) uniform,
        terminal_uniform=(T,B) uniform, optionally decoder=(B,C,T,H,W) normal for the decoded video) — so that seeded runs are comparable draw for draw across devices; by default the
        draws come from torch's CUDA generator in the reference's per-frame order (randn latent, rand terminal, rand per
        action type, randn context)."""
        if return_for_policy_optimization:            # reference dreamer4.py:6342-6347
            return_agent_actions = True
            return_log_probs_and_values = True
            return_rewards_per_frame = True
            return_terminals = return_terminals or self.predict_terminals
        for name, v in dict(prompt_proprio=prompt_proprio, prompt_continuous_actions=prompt_continuous_actions,
                            latent_gene_ids=latent_gene_ids).items():
            if exists(v):
                raise NotImplementedError(f'generate({name}=...): proprioception / continuous actions are "next" rows (SURVEY.md section 8f)')
        has_tokenizer = exists(self.video_tokenizer)
        return_decoded_video = default(return_decoded_video, has_tokenizer)          # reference dreamer4.py:6694-6695
        if return_decoded_video and not has_tokenizer:
            raise ValueError('return_decoded_video needs a video_tokenizer attached to the model')
        assert not (exists(prompt) and exists(prompt_latents)), 'cannot pass in both prompt video and prompt latents'
        if exists(prompt):                                                           # reference dreamer4.py:6377-6387
            assert has_tokenizer, 'a video prompt needs a video_tokenizer attached to the model'
            if prompt.ndim == 4:
                prompt = prompt[:, :, None]
            if prompt.shape[1] != self.video_tokenizer.channels:
                prompt = prompt.expand(-1, self.video_tokenizer.channels, -1, -1, -1)
            prompt_latents = self.video_tokenizer.tokenize(prompt)
        if not use_time_cache:
            raise NotImplementedError('use_time_cache=False: the native path always decodes over the in-place KV cache')
        assert math.log2(num_steps).is_integer(), f'number of steps {num_steps} must be a power of 2'
        assert 0 < num_steps <= self.max_steps
        if num_steps == 1:
            raise ValueError('num_steps=1 indexes step_size_embed out of range in the reference (SURVEY.md section 8a); use >= 2')
        c = self.cfg
        B, T, N, Dl, D = batch_size, time_steps, c.num_latent_tokens, c.dim_latent, c.dim
        dev = self.device
        f32 = dict(device=dev, dtype=torch.float32)

        # ---- prompt and resumed cache (reference dreamer4.py:6377-6402; env.py:464-484)
        P, prompt_lat = 0, None
        if exists(prompt_latents):
            prompt_lat = prompt_latents[:, :, 0] if prompt_latents.ndim == 5 else prompt_latents      # lone view dim (6394)
            assert prompt_lat.shape[0] == B and tuple(prompt_lat.shape[2:]) == (N, Dl), f'prompt_latents {tuple(prompt_latents.shape)}'
            prompt_lat = prompt_lat.to(**f32).contiguous()
            P = prompt_lat.shape[1]
            assert P < T, f'time_steps={T} must exceed the {P} prompt frames'
        resumed_kv = None
        off = 0               # frames in the cache that are not part of this call's output (a time cache without prompt latents)
        if exists(time_cache):
            resumed_kv = time_cache.main.next_kv_cache if exists(time_cache.main) else None
            cached = time_cache.main.token_count if exists(time_cache.main) else 0
            if P == 0:
                # the reference's tests/test_dreamer.py::test_cache_generate flow: time_steps NEW frames on top of the cached ones,
                # at cache positions cached .. cached + T - 1 (their rotary offset, dreamer4.py:3010); only the new frames come back
                off = cached
            else:
                assert cached == P, f'time_cache holds {cached} frames but the prompt has {P}: pass the latents of exactly the cached frames'
        lib, ctx, kv = self._adopt_time_cache(resumed_kv, P + off, B, T + off, agent_index, grow=P > 0 or exists(time_cache))
        L = c.num_time_layers
        self._kv_epoch += 1

        return_agent_actions = (return_agent_actions or return_log_probs_and_values) and c.has_actions
        want_heads = return_agent_actions
        should_term = return_terminals and self.predict_terminals
        if isinstance(tasks, int):
            tasks = torch.full((B,), tasks, device=dev, dtype=torch.long)
        if exists(tasks):
            assert tasks.shape[0] == B
            tasks = tasks.to(dev, torch.long).contiguous()

        A = c.total_actions
        na = len(c.num_discrete_actions)
        latents = torch.empty(B, T, N, Dl, **f32)
        agent_embed = torch.empty(B, T, D, **f32)
        rewards = torch.empty(B, T, **f32)
        values = torch.empty(B, T, **f32) if want_heads else None
        log_probs = torch.empty(B, T, na, **f32) if want_heads else None
        logits = torch.empty(B, T, A, **f32) if want_heads else None
        lens = torch.full((B,), T, device=dev, dtype=torch.long)
        terminals = torch.zeros(B, device=dev, dtype=torch.bool)
        term_u8 = terminals.view(torch.uint8)
        # The action history `decoded` that conditions frame t is decoded[:, :t] right-padded with index 0, shifted by one
        # frame, and a zero token when there is no history at all (reference dreamer4.py:6519-6522, 7111-7126): rows past
        # n_dec stay zero here, and frame t reads row t-1.  Sampled actions append at row n_dec (6645) - row t unless the
        # prompt's action count differs from its frame count.
        n_dec = 0
        actions = None
        if exists(prompt_discrete_actions):
            assert c.has_actions and prompt_discrete_actions.shape[0] == B
            pa = prompt_discrete_actions if prompt_discrete_actions.ndim == 3 else prompt_discrete_actions[..., None]
            assert pa.shape[-1] == na, f'prompt_discrete_actions {tuple(prompt_discrete_actions.shape)}'
            n_dec = pa.shape[1]
        if want_heads or n_dec > 0:
            actions = torch.zeros(B, max(T, n_dec + T - P), na, device=dev, dtype=torch.long)
            if n_dec > 0:
                actions[:, :n_dec] = pa.to(dev, torch.long)
        if P > 0:
            latents[:, :P] = prompt_lat.clamp(-1., 1.)                        # the final clamp covers the prompt too (6686)
            if exists(prompt_rewards):
                assert tuple(prompt_rewards.shape) == (B, P), f'prompt_rewards {tuple(prompt_rewards.shape)} != {(B, P)}'
                rewards[:, :P] = prompt_rewards.to(**f32)
            else:
                assert not (return_rewards_per_frame or return_agent_actions), \
                    'prompt_rewards (b, prompt frames) is required to return an Experience from a prompted rollout (reference :6741-6743)'

        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

        def prev_actions_of(t):
            if t == 0 or n_dec == 0:
                return None, 0
            return C.c_void_p(actions[:, t - 1].data_ptr()), actions.stride(0)

        if P > 0 and not exists(time_cache):
            # Cold prompt.  The reference re-runs its uncached multi-frame forward over [prompt, new frame] on every pass
            # of the first new frame (6529-6546).  Prompt frames sit at signal level max_steps-1 with context noise equal to
            # themselves (6400, 6497) and time attention is causal, so that is one clean pass per prompt frame appending
            # its keys/values - the same pass d4_frame ends every frame with.
            step_log2 = int(math.log2(self.max_steps // num_steps))
            scratch = torch.empty(B, D, **f32)
            for p in range(P):
                pa_ptr, pa_stride = prev_actions_of(p)
                frame = prompt_lat[:, p].contiguous()
                check(lib.d4_pass(ctx, B, ptr(frame), self.max_steps - 1, step_log2, pa_ptr, pa_stride, ptr(tasks), p, 1,
                                  None, ptr(scratch), stream))

        io = _lib.d4_frame_io()
        frames = P
        keep = []
        for t in range(P, T):
            if exists(noise):
                nl = noise['latent'][t]
                au = noise['action_uniform'][t] if want_heads else None
                tu = noise['terminal_uniform'][t] if should_term else None
            else:                                       # reference draw order: dreamer4.py:6475, 6611, 6637, 6670
                nl = torch.randn(B, N, Dl, **f32)
                tu = torch.rand(B, **f32) if should_term else None
                au = torch.cat([torch.rand(B, n, **f32) for n in c.num_discrete_actions], dim=-1) if want_heads else None
            nl = nl.to(**f32).contiguous()
            au = au.to(**f32).contiguous() if exists(au) else None
            tu = tu.to(**f32).contiguous() if exists(tu) else None
            keep = [nl, au, tu]
            io.noise_latent, io.action_uniform, io.terminal_uniform = ptr(nl), ptr(au), ptr(tu)
            io.prev_actions, io.pa_stride = prev_actions_of(t)
            io.tasks = ptr(tasks)
            io.latents, io.latents_bs = C.c_void_p(latents[:, t].data_ptr()), latents.stride(0)
            io.agent_embed, io.agent_bs = C.c_void_p(agent_embed[:, t].data_ptr()), agent_embed.stride(0)
            io.rewards, io.rewards_bs = C.c_void_p(rewards[:, t].data_ptr()), rewards.stride(0)
            if want_heads:
                io.values, io.values_bs = C.c_void_p(values[:, t].data_ptr()), values.stride(0)
                io.actions, io.actions_bs = C.c_void_p(actions[:, n_dec].data_ptr()), actions.stride(0)
                io.log_probs, io.log_probs_bs = C.c_void_p(log_probs[:, t].data_ptr()), log_probs.stride(0)
                io.logits, io.logits_bs = C.c_void_p(logits[:, t].data_ptr()), logits.stride(0)
            else:
                io.values = io.actions = io.log_probs = io.logits = None
            io.lens, io.terminals = ptr(lens), ptr(term_u8)
            check(lib.d4_frame(ctx, B, t + off, num_steps, float(discrete_temperature), C.byref(io), stream))
            if want_heads:
                n_dec += 1
            if not exists(noise) and context_signal_noise > 0.:
                torch.randn(B, N, Dl, **f32)            # dreamer4.py:6670: consumed by the reference, numerically dead with the KV cache
            frames = t + 1
            if should_term and bool(terminals.all()):   # reference dreamer4.py:6681
                break
        del keep

        Tg = frames
        latents = latents[:, :Tg]
        if off > 0 and should_term:
            lens = torch.where(terminals, lens - off, lens)      # the engine counts frames by cache position
        next_kv = None
        if L > 0:
            next_kv = kv[:L, :, :, :, :off + Tg]
            next_kv._d4_epoch = self._kv_epoch          # lets a resumed call recognise this view as current (see above)
        tc = DynamicsIntermediates(main=TransformerIntermediates(next_kv_cache=next_kv, token_count=off + Tg))
        video = None
        if return_decoded_video:                                                     # reference dreamer4.py:6699-6711
            dec_kw = dict(noise=noise['decoder']) if (exists(noise) and 'decoder' in noise) else {}      # injected start noise (tests)
            video = self.video_tokenizer.decode(latents, height=image_height, width=image_width, **dec_kw)
        if not (return_rewards_per_frame or return_agent_actions):
            out = video if return_decoded_video else latents
            return (out, tc) if return_time_cache else out

        # prompt frames carry latents, rewards and actions only; everything decoded off the agent token covers the new
        # frames (reference dreamer4.py:6620-6662 accumulate from empty)
        rewards = rewards[:, :Tg]
        step_mask = (torch.arange(Tg, device=dev)[None, :] < lens[:, None]).float()
        gen = Experience(
            latents=latents,
            video=video,
            agent_embed=agent_embed[:, P:Tg] if store_agent_embed else None,
            old_action_unembeds=Actions(logits[:, P:Tg], None) if (want_heads and store_old_action_unembeds) else None,
            step_size=self.max_steps // num_steps, agent_index=agent_index, lens=lens, is_truncated=~terminals, terminals=terminals,
            is_from_world_model=True,
            episode_return=(rewards * step_mask).sum(dim=-1),
            rewards=rewards if return_rewards_per_frame else None,
            actions=Actions(actions[:, :n_dec], None) if return_agent_actions else None,
            log_probs=Actions(log_probs[:, P:Tg], None) if (return_log_probs_and_values and want_heads) else None,
            values=values[:, P:Tg] if (return_log_probs_and_values and want_heads) else None)
        return (gen, tc) if return_time_cache else gen

    # ------------------------------------------------------------------ interact_with_env

    @torch.no_grad()
    def interact_with_env(self, env, seed=None, agent_index=0, num_steps=4, max_timesteps=16, env_is_vectorized=False,
                          use_time_cache=True, store_agent_embed=True, store_old_action_unembeds=True, obs_to_latents_fn=None):
        """One real-environment episode (batch of episodes when vectorized) driven by the policy head (reference
        dreamer4.py:5470-5889).  Per env step: `obs_to_latents_fn(self, obs, cache) -> (latents (b, 1, n, d), cache)`, one
        native `d4_observe` (the clean pass over the time cache conditioned on the previous action, value head, policy head,
        gumbel-argmax sample, log-prob), `env.step`.  Episodes cut by `max_timesteps` get the reference's bootstrap step: the
        final observation is evaluated once more for its value and every per-step record is right-padded by one (:5790-5853).

        Without a VideoTokenizer on this path `obs_to_latents_fn` is required; it is called with the same
        (self, obs, cache) arguments at the bootstrap step (the reference drops `self` there, :5793)."""
        if not exists(obs_to_latents_fn):
            if not exists(self.video_tokenizer):
                raise NotImplementedError('interact_with_env needs obs_to_latents_fn or an attached video_tokenizer: state observations '
                                          'go through state_to_latents, a "next" row (SURVEY.md section 8f)')

            def obs_to_latents_fn(model, obs, cache):       # reference dreamer4.py:5588: one frame through the encoder's time cache
                image = obs['image'] if isinstance(obs, dict) else obs
                image = torch.as_tensor(image, dtype=torch.float32).to(model.device)
                if not env_is_vectorized:
                    image = image[None]
                return model.video_tokenizer(image[:, :, None], return_latents=True, time_cache=cache, return_time_cache=True)
        if not use_time_cache:
            raise NotImplementedError('use_time_cache=False: the native path always runs over the in-place KV cache')
        c = self.cfg
        assert c.has_actions, 'interact_with_env needs a model with discrete actions'
        assert self.max_steps % num_steps == 0
        dev = self.device
        N, Dl, D, A, na = c.num_latent_tokens, c.dim_latent, c.dim, c.total_actions, len(c.num_discrete_actions)
        f32 = dict(device=dev, dtype=torch.float32)
        as_f32 = lambda v: torch.as_tensor(v, dtype=torch.float32).to(dev)

        def as_obs(obs):                                      # reference :5513-5521, :5712-5720
            if isinstance(obs, dict):
                return obs
            obs = obs if torch.is_tensor(obs) else as_f32(obs)
            return dict(image=obs) if obs.ndim >= 3 else dict(state=obs)

        def frames_of(obs):                                   # (image (b c 1 h w) | None, state (b d) | None), reference :5528-5541
            image = as_f32(obs['image']) if 'image' in obs else None
            state = as_f32(obs['state']) if 'state' in obs else None
            if not env_is_vectorized:
                image, state = (t[None] if exists(t) else None for t in (image, state))
            return (image[:, :, None] if exists(image) else None), state

        obs = env.reset(seed=seed)
        obs = as_obs(obs[0] if isinstance(obs, tuple) else obs)
        assert 'image' in obs or 'state' in obs
        image, state = frames_of(obs)
        B = image.shape[0] if exists(image) else state.shape[0]
        video_frames, states = ([image] if exists(image) else []), ([state] if exists(state) else [])

        T = max_timesteps + 1                                 # + the bootstrap frame
        lib, ctx = self._engine(B, T, agent_index, grow=True)
        self._kv_epoch += 1
        latents = torch.zeros(B, T, N, Dl, **f32)
        agent_embed = torch.zeros(B, T, D, **f32)
        values = torch.zeros(B, T, **f32)
        rewards = torch.zeros(B, T, **f32)
        actions = torch.zeros(B, T, na, device=dev, dtype=torch.long)
        log_probs = torch.zeros(B, T, na, **f32)
        logits = torch.zeros(B, T, A, **f32)
        scratch = torch.empty(B, N, Dl, **f32)
        terminated_any = torch.zeros(B, dtype=torch.bool, device=dev)
        truncated_any = torch.zeros(B, dtype=torch.bool, device=dev)
        done = torch.zeros(B, dtype=torch.bool, device=dev)
        lens = torch.zeros(B, dtype=torch.long, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        io = _lib.d4_frame_io()

        def observe(t, obs, cache, uniform):
            """latents of `obs` -> frame t of the records; returns the tokenizer-side cache."""
            lat, cache = obs_to_latents_fn(self, obs, cache)
            lat = lat.to(**f32).reshape(B, N, Dl).contiguous()
            latents[:, t] = lat
            io.noise_latent, io.action_uniform, io.terminal_uniform = ptr(lat), ptr(uniform), None
            io.prev_actions, io.pa_stride = (C.c_void_p(actions[:, t - 1].data_ptr()), actions.stride(0)) if t > 0 else (None, 0)
            io.tasks = None
            io.latents, io.latents_bs = ptr(scratch), scratch.stride(0)
            io.agent_embed, io.agent_bs = C.c_void_p(agent_embed[:, t].data_ptr()), agent_embed.stride(0)
            io.rewards, io.rewards_bs = None, 0
            io.values, io.values_bs = C.c_void_p(values[:, t].data_ptr()), values.stride(0)
            io.actions, io.actions_bs = C.c_void_p(actions[:, t].data_ptr()), actions.stride(0)
            io.log_probs, io.log_probs_bs = C.c_void_p(log_probs[:, t].data_ptr()), log_probs.stride(0)
            io.logits, io.logits_bs = C.c_void_p(logits[:, t].data_ptr()), logits.stride(0)
            io.lens = io.terminals = None
            check(lib.d4_observe(ctx, B, t, num_steps, 1., C.byref(io), stream))
            return cache

        cache = None
        step = 0
        while not bool(done.all()):
            step += 1
            t = step - 1
            uniform = torch.cat([torch.rand(B, n, **f32) for n in c.num_discrete_actions], dim=-1)      # the sampler's draws (:5657)
            cache = observe(t, obs, cache, uniform)
            act = actions[:, t].cpu().numpy()                 # (b, na); the reference's host sync, :5679-5690
            if not env_is_vectorized:
                act = act[0]
                if act.size == 1:
                    act = int(act.item())
            out = env.step(act)
            assert 2 <= len(out) <= 5, f'env.step returned {len(out)} values'
            next_obs, reward = out[0], out[1]
            flag = lambda i: torch.as_tensor(out[i]).to(dev).reshape(B).bool() if len(out) > i else torch.zeros(B, dtype=torch.bool, device=dev)
            terminated, truncated = flag(2), flag(3)
            lens = torch.where(done, lens, lens + 1)          # :5726
            terminated_any |= terminated
            truncated_any |= truncated
            if step >= max_timesteps:
                truncated_any |= ~terminated_any              # :5731-5732
            done |= terminated_any | truncated_any
            rewards[:, t] = as_f32(reward).reshape(B)
            obs = as_obs(next_obs)
            assert 'image' in obs or 'state' in obs
            image, state = frames_of(obs)
            if exists(state):
                states.append(state)
            if exists(image):
                video_frames.append(image)
            need_bootstrap = truncated_any & ~terminated_any
            if bool(done.all()) and bool(need_bootstrap.any()):                                           # :5790-5853
                # value of the state the episode was cut at; its reward / action / log-prob rows stay zero (the right-pad)
                observe(step, obs, cache, torch.full((B, A), 0.5, **f32))
                actions[:, step] = 0
                log_probs[:, step] = 0.
                lens = torch.where(need_bootstrap, lens + 1, lens)
                step += 1
                break

        Tg = step
        step_mask = (torch.arange(Tg, device=dev)[None, :] < lens[:, None]).float()
        return Experience(
            latents=latents[:, :Tg],
            video=torch.cat(video_frames, dim=2)[:, :, :Tg] if video_frames else None,
            critic_state=torch.stack(states, dim=1)[:, :Tg] if states else None,
            rewards=rewards[:, :Tg], actions=Actions(actions[:, :Tg], None), log_probs=Actions(log_probs[:, :Tg], None),
            values=values[:, :Tg],
            old_action_unembeds=Actions(logits[:, :Tg], None) if store_old_action_unembeds else None,
            agent_embed=agent_embed[:, :Tg] if store_agent_embed else None,
            step_size=self.max_steps // num_steps, agent_index=agent_index, is_truncated=truncated_any, terminals=terminated_any,
            lens=lens, is_from_world_model=False, episode_return=(rewards[:, :Tg] * step_mask).sum(dim=-1))

    # ------------------------------------------------------------------ learn_from_experience

    def learn_from_experience(self, experience: Experience, policy_optim=None, value_optim=None, only_learn_policy_value_heads=True,
                              objective='ppo', use_delight_gating=None, delight_temperature=None, normalize_advantages=None, eps=1e-6):
        """Actor/critic losses with gradients (reference dreamer4.py:5893-6305).  Both losses and every head-parameter
        gradient are produced by one native call; the returned scalars carry an autograd node that deposits those
        gradients on `.backward()` exactly like the reference's graph would."""
        if objective not in _lib.OBJECTIVES:
            raise ValueError(f'unknown objective {objective}')              # reference dreamer4.py:6214-6215
        if not only_learn_policy_value_heads:
            raise NotImplementedError('only_learn_policy_value_heads=False needs the world-model backward (outside this path)')
        c = self.cfg
        assert c.has_actions, 'learn_from_experience needs a model with discrete actions'
        exp = experience
        assert isinstance(exp, Experience)
        assert exists(exp.agent_embed), 'the native path learns from stored agent embeds (generate(store_agent_embed=True))'
        assert all(map(exists, (exp.log_probs, exp.actions, exp.values, exp.rewards, exp.step_size))), \
            'the generations need to contain the log probs, values, and rewards for policy optimization - world_model.generate(..., return_log_probs_and_values = True)'
        assert not (exists(exp.actions.continuous) and exp.actions.continuous.numel() > 0), 'continuous actions are a "next" row (SURVEY.md section 8f)'
        B, T = exp.latents.shape[:2]
        dev = self.device
        na, A = len(c.num_discrete_actions), c.total_actions
        disc = exp.actions.discrete if exp.actions.discrete.ndim == 3 else exp.actions.discrete[..., None]        # reference :6037-6038
        old_disc = exp.log_probs.discrete if exp.log_probs.discrete.ndim == 3 else exp.log_probs.discrete[..., None]
        # every per-step record must cover the same (B, T) steps as the latents: the native call reads B*T rows from each.  (A
        # prompted generate(..., return_for_policy_optimization=True) returns latents / rewards over prompt + new frames but
        # agent embeds, values and log-probs over the new frames only; the reference fails on that in calc_gae - slice the
        # Experience to the new frames before learning from it.)
        for name, t, shape in (('agent_embed', exp.agent_embed, (B, T, c.dim)), ('rewards', exp.rewards, (B, T)), ('values', exp.values, (B, T)),
                               ('actions.discrete', disc, (B, T, na)), ('log_probs.discrete', old_disc, (B, T, na))):
            if tuple(t.shape) != shape:
                raise ValueError(f'learn_from_experience: experience.{name} has shape {tuple(t.shape)}, expected {shape} to match latents {tuple(exp.latents.shape)}')
        if exists(exp.lens) and tuple(exp.lens.shape) != (B,):
            raise ValueError(f'learn_from_experience: experience.lens has shape {tuple(exp.lens.shape)}, expected {(B,)}')
        lib, ctx = self._engine(*(self._ctx_key[:2] if self._ctx_key else (B, T)), default(exp.agent_index, 0))
        use_gate = default(use_delight_gating, self.use_delight_gating)
        temp = default(delight_temperature, self.delight_temperature)
        norm_adv = default(default(normalize_advantages, self.normalize_advantages), objective != 'pmpo')    # reference :6021

        f32 = dict(device=dev, dtype=torch.float32)
        cont = lambda t, dt: t.detach().to(device=dev, dtype=dt).contiguous()
        agent = cont(exp.agent_embed, torch.float32)
        rewards, values = cont(exp.rewards, torch.float32), cont(exp.values, torch.float32)
        actions, old_lp = cont(disc, torch.long), cont(old_disc, torch.float32)
        lens = cont(default(exp.lens, torch.full((B,), T, device=dev)), torch.long)
        # a missing is_truncated means "every episode was cut": the last step is bootstrap-only (reference :5943-5944, :5954)
        is_trunc = cont(default(exp.is_truncated, torch.ones(B, dtype=torch.bool, device=dev)), torch.bool)
        if tuple(is_trunc.shape) != (B,):
            raise ValueError(f'learn_from_experience: experience.is_truncated has shape {tuple(is_trunc.shape)}, expected {(B,)}')
        support, _ = hl_gauss_tables(*c.value_range, c.value_num_bins, dev)
        sigma_sqrt2 = math.sqrt(2.) * c.hl_gauss_sigma_to_bin_ratio * (c.value_range[1] - c.value_range[0]) / c.value_num_bins

        params = dict(self.named_parameters())
        pol_names = [k for _, k in mlp_param_names('policy_head', c.policy_head_mlp_depth + 2)]
        val_names = [k for _, k in mlp_param_names('value_head', c.value_head_mlp_depth + 2)]
        un_name = 'action_embedder.discrete_action_unembed'
        grads = {k: torch.zeros_like(params[k], dtype=torch.float32) for k in pol_names + val_names + [un_name]}
        losses = torch.zeros(4, **f32)
        returns = torch.empty(B, T, **f32)
        adv = torch.empty(B, T, **f32)

        io = _lib.d4_learn_io()
        io.B, io.T = B, T
        io.agent_embed, io.rewards, io.old_values = ptr(agent), ptr(rewards), ptr(values)
        io.actions, io.old_log_probs, io.lens = ptr(actions), ptr(old_lp), ptr(lens)
        io.is_truncated = ptr(is_trunc.view(torch.uint8))
        io.terminals = None
        io.gamma, io.lam, io.eps_clip = self.gae_discount_factor, self.gae_lambda, self.ppo_eps_clip
        io.entropy_weight, io.delight_temperature, io.zscore_eps = self.policy_entropy_weight, temp, eps
        io.use_delight_gating, io.normalize_advantages = int(bool(use_gate)), int(bool(norm_adv))
        io.value_support = ptr(support)
        io.value_sigma_sqrt2, io.hl_eps = sigma_sqrt2, c.hl_gauss_eps
        io.value_lo, io.value_hi = c.value_range
        io.losses, io.returns, io.advantages = ptr(losses), ptr(returns), ptr(adv)
        io.objective = _lib.OBJECTIVES[objective]
        keep = []
        if self.keep_reward_ema_stats:                                      # reference dreamer4.py:5987-6013
            # the running return statistics need this batch's lambda-returns first: the native scan (d4_gae) on the masked rewards /
            # values, then the quantile-filtered mean / variance and their EMA - a handful of torch ops over (B, T) scalars,
            # plumbing like the optimizer step - and d4_learn normalises returns and old values by them before subtracting
            steps = torch.arange(T, device=dev)[None, :]
            in_len = steps < lens[:, None]
            learn_mask = steps < (lens - is_trunc.long())[:, None]
            gae_mask = steps < (lens - 1).clamp(min=0)[:, None]
            r_m, v_m = (rewards * in_len).contiguous(), (values * in_len).contiguous()
            gm_u8, lm_u8 = gae_mask.to(torch.uint8).contiguous(), learn_mask.to(torch.uint8).contiguous()      # named: alive across the call
            stream0 = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            check(lib.d4_gae(B, T, ptr(r_m), ptr(v_m), ptr(gm_u8), ptr(lm_u8), self.gae_discount_factor, self.gae_lambda, ptr(returns), stream0))
            with torch.no_grad():
                rs = returns[learn_mask]
                lo, hi = torch.quantile(rs, self.reward_quantile_filter.to(dev)).unbind()
                rs = torch.minimum(torch.maximum(rs, lo), hi)
                decay = 1. - self.reward_ema_decay
                self.ema_returns_mean.lerp_(rs.mean(), decay)
                self.ema_returns_var.lerp_(rs.var(correction=0), decay)
                ema = torch.stack((self.ema_returns_mean, self.ema_returns_var.clamp(min=1e-5).sqrt())).to(**f32).contiguous()
            keep.append(ema)
            io.returns_ema = ptr(ema)
        if objective == 'pmpo':                                             # reference dreamer4.py:6127-6182
            io.pmpo_pos_to_neg_weight, io.pmpo_kl_div_loss_weight = self.pmpo_pos_to_neg_weight, self.pmpo_kl_div_loss_weight
            io.pmpo_reverse_kl = int(bool(self.pmpo_reverse_kl))
            if self.pmpo_kl_div_loss_weight > 0.:
                old = exp.old_action_unembeds
                assert exists(old) and exists(old.discrete), 'pmpo with a KL weight needs generate(store_old_action_unembeds=True)'
                old_logits = cont(old.discrete, torch.float32)
                assert old_logits.shape == (B, T, sum(c.num_discrete_actions)), f'old_action_unembeds {tuple(old_logits.shape)}'
                io.old_action_unembeds, io.old_action_unembeds_ld = ptr(old_logits), old_logits.stride(1)

        def fill(head, nl, arrs):
            for l in range(nl):
                arrs[0][l] = grads[f'{head}.layers.{l}.0.weight'].data_ptr()
                arrs[1][l] = grads[f'{head}.layers.{l}.0.bias'].data_ptr()
                if l < nl - 1:
                    arrs[2][l] = grads[f'{head}.layers.{l}.1.weight'].data_ptr()
                    arrs[3][l] = grads[f'{head}.layers.{l}.1.bias'].data_ptr()
        fill('policy_head', c.policy_head_mlp_depth + 2, (io.grad_policy_w, io.grad_policy_b, io.grad_policy_lnw, io.grad_policy_lnb))
        fill('value_head', c.value_head_mlp_depth + 2, (io.grad_value_w, io.grad_value_b, io.grad_value_lnw, io.grad_value_lnb))
        io.grad_unembed, io.grad_unembed_ld = ptr(grads[un_name]), grads[un_name].stride(0)

        ws_bytes = lib.d4_learn_workspace_bytes(ctx, B, T)
        ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=dev)
        ws = ws[(-ws.data_ptr()) % 256:][:ws_bytes]          # 256-byte aligned (a no-op offset for CUDA allocations)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        check(lib.d4_learn(ctx, C.byref(io), ptr(ws), ws_bytes, stream))
        ws.record_stream(torch.cuda.current_stream(dev))

        self.last_learn_aux = dict(returns=returns, advantages=adv, policy_surrogate=losses[2], entropy_term=losses[3])
        pol_params = [params[k] for k in pol_names + [un_name]]
        val_params = [params[k] for k in val_names]
        policy_loss = _DepositGrads.apply(losses[0].clone(), [grads[k] for k in pol_names + [un_name]], *pol_params)
        value_loss = _DepositGrads.apply(losses[1].clone(), [grads[k] for k in val_names], *val_params)

        if exists(policy_optim):          # reference dreamer4.py:6246-6250
            policy_loss.backward()
            policy_optim.step()
            policy_optim.zero_grad()
        if exists(value_optim):           # reference dreamer4.py:6299-6303
            value_loss.backward()
            value_optim.step()
            value_optim.zero_grad()
        return policy_loss, value_loss


class _DepositGrads(torch.autograd.Function):
    """A scalar loss whose gradients with respect to `params` were already computed natively."""

    @staticmethod
    def forward(ctx, loss, grads, *params):
        ctx.grads = grads
        return loss.clone()

    @staticmethod
    def backward(ctx, grad_out):
        return (None, None, *[g * grad_out for g in ctx.grads])
