"""CPU oracle for the Dreamer-4 imagination hot path.  TEST INFRASTRUCTURE — NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this file.  Nothing under `dreamer4_b200/` does, and the product path raises
when its CUDA extension is missing rather than falling back to anything here.

What it is: a plain fp32 PyTorch restatement of the reference's algorithm for
`DynamicsWorldModel.generate` + `learn_from_experience` (default branches), written as pure
functions over a reference-layout `state_dict` so that it shares no code with the product.
Every function cites the reference lines it follows (paths relative to /root/reference,
D4 = dreamer4/dreamer4.py).

Parity status
  * In-tree arithmetic (attention, transformer, token assembly, flow step, generate loop,
    GAE/PPO/value loss): PINNED — `tests/test_oracle_golden.py` checks this file against golden
    vectors produced by executing the reference's own dreamer4.py (oracle/make_golden.py).
  * Arithmetic inside un-vendored third-party packages (x-mlps `create_mlp` / `Ensemble`,
    hl-gauss `HLGaussLoss`, `MultiCategorical`, `AssocScan`, `masked_mean`): PARITY UNPINNED —
    the packages are absent from /root/reference and from this image; they are restated from
    their published behaviour (oracle/shims/*) and the reference's tests hold no golden vectors
    for them.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# configuration


@dataclass
class OracleConfig:
    dim: int
    dim_latent: int
    num_latent_tokens: int
    depth: int = 4
    num_spatial_tokens: int = 4
    num_register_tokens: int = 8
    time_block_every: int = 4
    attn_heads: int = 8
    attn_dim_head: int = 64
    query_heads: Optional[int] = None          # attn_kwargs['query_heads'] (GQA)
    attn_softclamp_value: float = 50.0
    max_steps: int = 64
    num_discrete_actions: tuple = ()
    multi_token_pred_len: int = 8
    policy_head_mlp_depth: int = 3
    value_head_mlp_depth: int = 3
    ff_activation: str = 'silu'                # 'silu' | 'gelu' (both gated, D4:2093)
    ff_expansion_factor: float = 4.0
    head_activation: str = 'silu'
    reward_range: tuple = (-20.0, 20.0)
    reward_num_bins: int = 255
    value_range: tuple = (-20.0, 20.0)
    value_num_bins: int = 255
    hl_gauss_sigma_to_bin_ratio: float = 2.0
    hl_gauss_eps: float = 1e-10
    predict_terminals: bool = False
    terminal_mlp_depth: int = 1
    num_tasks: int = 0
    pool_heads: int = 4
    pool_dim_head: int = 64
    gae_discount_factor: float = 0.997
    gae_lambda: float = 0.95
    ppo_eps_clip: float = 0.2
    policy_entropy_weight: float = 0.01
    use_delight_gating: bool = True
    delight_temperature: float = 1.0
    pmpo_pos_to_neg_weight: float = 0.5
    pmpo_reverse_kl: bool = True
    pmpo_kl_div_loss_weight: float = 0.3

    def __post_init__(self):
        nda = self.num_discrete_actions
        if isinstance(nda, int):
            nda = (nda,)
        self.num_discrete_actions = tuple(int(n) for n in nda if n > 0)
        if self.query_heads is None:
            self.query_heads = self.attn_heads

    @property
    def has_actions(self):
        return len(self.num_discrete_actions) > 0

    @property
    def same_len(self):
        return self.num_spatial_tokens == self.num_latent_tokens

    @property
    def tokens_per_frame(self):        # D4:7222
        return 1 + self.num_spatial_tokens + self.num_register_tokens + int(self.has_actions) + 1

    @property
    def is_time(self):                 # D4:2845
        return [((i + 1) % self.time_block_every) == 0 for i in range(self.depth)]

    @property
    def ff_inner(self):                # D4:2094
        return int(self.dim * self.ff_expansion_factor * 2 / 3)


# --------------------------------------------------------------------------------------
# small ops


def rmsnorm(x, weight):
    """nn.RMSNorm(dim) with eps=None -> finfo.eps (D4:1906, 2089, 2822)."""
    return F.rms_norm(x, (x.shape[-1],), weight, None)


def l2norm(t):
    """D4:521-522."""
    return F.normalize(t, dim=-1, p=2)


def multi_head_rmsnorm(x, gamma):
    """D4:1663-1679.  x (b h n d), gamma (h d)."""
    d = x.shape[-1]
    return l2norm(x) * ((gamma + 1.0) * d ** 0.5)[:, None, :]


def rotary_angles(inv_freq, seq_len, offset):
    """D4:1614-1624."""
    t = torch.arange(seq_len, dtype=inv_freq.dtype) + offset
    freqs = t[:, None] * inv_freq[None, :]
    return torch.cat((freqs, freqs), dim=-1)


def apply_rotations(rot, t):
    """D4:1626-1659 (non-GQA-rotation branch: rot is (n d))."""
    x1, x2 = t.chunk(2, dim=-1)
    return t * rot.cos() + torch.cat((-x2, x1), dim=-1) * rot.sin()


def softclamp(t, value):
    """D4:527-528."""
    return (t / value).tanh() * value


def activation_fn(name):
    return dict(silu=F.silu, gelu=F.gelu, relu=F.relu)[name]


def naive_attend(q, k, v, softclamp_value=None, mask=None):
    """D4:1683-1756.  q (b hq i d), k/v (b hk j d).  The causal mask is omitted: on this path the
    time-attention query length is 1 with all keys in the past (D4:1738 yields an empty mask)."""
    groups = q.shape[1] // k.shape[1]
    if softclamp_value is None and mask is None:
        # SDPA branch D4:1694-1701 (used by the pools and the final agent cross-attention)
        if groups > 1:
            k = k.repeat_interleave(groups, dim=1)
            v = v.repeat_interleave(groups, dim=1)
        sim = torch.einsum('bhid,bhjd->bhij', q, k) * q.shape[-1] ** -0.5
        return torch.einsum('bhij,bhjd->bhid', sim.softmax(dim=-1), v)
    b, hq, i, d = q.shape
    q = q.reshape(b, hq // groups, groups, i, d)
    sim = torch.einsum('bhgid,bhjd->bhgij', q, k) * d ** -0.5
    if softclamp_value is not None:
        sim = softclamp(sim, softclamp_value)
    if mask is not None:
        sim = sim.masked_fill(~mask, -torch.finfo(sim.dtype).max)
    out = torch.einsum('bhgij,bhjd->bhgid', sim.softmax(dim=-1), v)
    return out.reshape(b, hq, i, d)


def space_mask(seq_len, num_special):
    """D4:1769-1783, 1857-1861: non-special queries may not attend to special (agent) keys."""
    qi = torch.arange(seq_len)[:, None]
    kj = torch.arange(seq_len)[None, :]
    start = seq_len - num_special
    return ~((qi < start) & (kj >= start))


# --------------------------------------------------------------------------------------
# blocks (functional, over a state-dict prefix)


def _split_heads(x, d):
    b, n, _ = x.shape
    return x.reshape(b, n, -1, d).transpose(1, 2)


def attention(sd, p, tokens, d_head, *, context=None, kv_cache=None, rot=None, residual_values=None,
              softclamp_value=None, mask=None, belief=True):
    """Attention.forward, D4:1968-2075.  tokens (b n D).  Returns (out, (k, v)) with k, v the
    concatenated keys/values this call attended over (D4:2075)."""
    x = rmsnorm(tokens, sd[p + 'norm.weight'])
    q = x @ sd[p + 'to_q.weight'].T
    has_context = context is not None
    if has_context:
        ctx = rmsnorm(context, sd[p + 'norm_context.weight']) if (p + 'norm_context.weight') in sd else context
    else:
        ctx = x
    k = ctx @ sd[p + 'to_k.weight'].T
    v = ctx @ sd[p + 'to_v.weight'].T
    q, k, v = (_split_heads(t, d_head) for t in (q, k, v))
    if residual_values is not None:                                   # D4:2005-2012
        w, b_ = sd[p + 'to_learned_value_residual_mix.0.weight'], sd[p + 'to_learned_value_residual_mix.0.bias']
        mix = torch.sigmoid(x @ w.T + b_).transpose(1, 2)[..., None]  # (b h n 1)
        v = v.lerp(residual_values, mix)
    k = multi_head_rmsnorm(k, sd[p + 'k_heads_rmsnorm.gamma'])        # D4:2016-2017 (keys only)
    if rot is not None:                                               # D4:2021-2023
        q, k = apply_rotations(rot, q), apply_rotations(rot, k)
    v_for_belief = v
    if kv_cache is not None:                                          # D4:2032-2035
        ck, cv = kv_cache
        k, v = torch.cat((ck, k), dim=-2), torch.cat((cv, v), dim=-2)
    out = naive_attend(q, k, v, softclamp_value=softclamp_value, mask=mask)
    if belief and not has_context:                                    # D4:2049-2054
        vn = l2norm(v_for_belief)
        groups = q.shape[1] // vn.shape[1]
        if groups > 1:
            vn = vn.repeat_interleave(groups, dim=1)
        out = out - (out * vn).sum(dim=-1, keepdim=True) * vn
    gates = torch.sigmoid(x @ sd[p + 'to_gates.0.weight'].T).transpose(1, 2)[..., None]   # D4:2058-2060
    out = out * gates
    out = out.transpose(1, 2).reshape(tokens.shape[0], tokens.shape[1], -1)
    return out @ sd[p + 'to_out.weight'].T, (k, v)


def feedforward(sd, p, x, act):
    """FeedForward.forward (gated), D4:2105-2116."""
    x = rmsnorm(x, sd[p + 'norm.weight'])
    x = x @ sd[p + 'proj_in.weight'].T + sd[p + 'proj_in.bias']
    x, gates = x.chunk(2, dim=-1)
    x = x * activation_fn(act)(gates)
    return x @ sd[p + 'proj_out.weight'].T + sd[p + 'proj_out.bias']


def attention_pool(sd, p, x, hiddens, cfg):
    """Residual(AttentionPool), D4:2143-2177, 1869-1883.  x (b s D), hiddens list of (b s D)."""
    b, s, D = x.shape
    ctx = torch.stack(hiddens, dim=-2).reshape(b * s, len(hiddens), D)
    out, _ = attention(sd, p + 'fn.attn.', x.reshape(b * s, 1, D), cfg.pool_dim_head, context=ctx, belief=False)
    return x + out.reshape(b, s, D)


def learned_queries_pool(sd, p, x, d_head):
    """LearnedQueriesAttentionPool.forward, D4:2203-2210.  x (b n d_in) -> (b nq D)."""
    queries = sd[p + 'queries'][None].expand(x.shape[0], -1, -1)
    out, _ = attention(sd, p + 'attn.', queries, d_head, context=x, belief=False)
    return out


def mlp(sd, p, x, act):
    """x-mlps normed MLP as restated in oracle/shims/x_mlps_pytorch/normed_mlp.py (PARITY UNPINNED):
    Linear -> LayerNorm -> act for every layer but the last, last layer a bare Linear."""
    i = 0
    while (p + f'layers.{i}.0.weight') in sd:
        x = x @ sd[p + f'layers.{i}.0.weight'].T + sd[p + f'layers.{i}.0.bias']
        if (p + f'layers.{i}.1.weight') in sd:
            x = F.layer_norm(x, (x.shape[-1],), sd[p + f'layers.{i}.1.weight'], sd[p + f'layers.{i}.1.bias'])
            x = activation_fn(act)(x)
        i += 1
    return x


class HLGauss:
    """hl-gauss-pytorch HLGaussLoss as restated in oracle/shims/hl_gauss_pytorch.py (PARITY UNPINNED);
    wrapper semantics D4:1041-1105 (clamp_to_range=True, sigma = 2 bin widths)."""

    def __init__(self, vrange, num_bins, ratio=2.0, eps=1e-10):
        lo, hi = vrange
        self.lo, self.hi, self.num_bins, self.eps = lo, hi, num_bins, eps
        self.support = torch.linspace(lo, hi, num_bins + 1).float()
        self.centers = (self.support[:-1] + self.support[1:]) / 2
        self.sigma_sqrt2 = math.sqrt(2.0) * ratio * (hi - lo) / num_bins

    def to_probs(self, target):
        target = target.clamp(self.lo, self.hi)
        cdf = torch.special.erf((self.support - target[..., None]) / self.sigma_sqrt2)
        z = cdf[..., -1] - cdf[..., 0]
        return (cdf[..., 1:] - cdf[..., :-1]) / z.clamp(min=self.eps)[..., None]

    def from_logits(self, logits):
        return (logits.softmax(dim=-1) * self.centers).sum(dim=-1)


# --------------------------------------------------------------------------------------
# the transformer pass for ONE new frame over a time-KV cache


def transformer_step(sd, cfg, tokens, kv_cache, token_count, prefix='transformer.', num_special=1, final_norm=False):
    """AxialSpaceTimeTransformer.forward for a single new frame (time == 1 after the slice at
    D4:2960-2961), default branches.  tokens (b S D); kv_cache list over time layers of (k, v)
    each (b*S, h, t, d) or None; returns (out tokens (b S D), new per-time-layer (k, v)).
    `cfg` supplies is_time / depth / attn_heads / attn_dim_head / attn_softclamp_value / ff_activation / pool_dim_head.
    The dynamics model's transformer has one special (agent) token and no final norm (D4:5183); the video tokenizer's
    encoder / decoder pass their own prefix, special-token count and final_norm=True (oracle/tokenizer_oracle.py)."""
    p = prefix
    b, S, D = tokens.shape
    d, h = cfg.attn_dim_head, cfg.attn_heads
    rot = rotary_angles(sd[p + 'time_rotary.inv_freq'], 1, token_count)                    # D4:3010
    mask = space_mask(S, num_special)                                                        # D4:2971-2973
    # value residual D4:3026-3027
    v0 = rmsnorm(tokens, sd[p + 'to_value_residual.0.weight']) @ sd[p + 'to_value_residual.1.weight'].T
    v0 = v0.reshape(b, S, h, d)
    layer_hiddens = [tokens]
    new_cache = []
    ti = 0
    for i, is_time in enumerate(cfg.is_time):
        ap = p + f'layers.{i}.2.fn.'
        if is_time:
            # 'b t s d -> b s t d' then pack '* t d' (D4:2848, 1981): batch = b*S, seq = 1
            x = tokens.reshape(b * S, 1, D)
            rv = v0.reshape(b * S, 1, h, d).transpose(1, 2)                                 # (bS h 1 d)
            cache = kv_cache[ti] if kv_cache is not None else None
            out, kv = attention(sd, ap, x, d, kv_cache=cache, rot=rot, residual_values=rv,
                                softclamp_value=cfg.attn_softclamp_value)
            new_cache.append(kv)
            ti += 1
            tokens = tokens + out.reshape(b, S, D)
        else:
            rv = v0.transpose(1, 2)                                                          # (b h S d)
            out, _ = attention(sd, ap, tokens, d, residual_values=rv,
                               softclamp_value=cfg.attn_softclamp_value, mask=mask)
            tokens = tokens + out
        layer_hiddens.append(tokens)                                                         # D4:3172
        tokens = tokens + feedforward(sd, p + f'layers.{i}.3.fn.', tokens, cfg.ff_activation)
        layer_hiddens.append(tokens)                                                         # D4:3216
        if i != cfg.depth - 1:                                                               # D4:2875, 3222
            tokens = attention_pool(sd, p + f'attn_pools.{i}.', tokens, layer_hiddens, cfg)
    # final agent cross attention + ff, D4:3227-3238 (SDPA branch: no softclamp, no mask, no belief)
    non_special, special = tokens[:, :-num_special], tokens[:, -num_special:]
    out, _ = attention(sd, p + 'final_special_cross_attn.fn.', special, d, context=non_special, belief=False)
    special = special + out
    special = special + feedforward(sd, p + 'final_special_ff.fn.', special, cfg.ff_activation)
    tokens = torch.cat((non_special, special), dim=1)
    tokens = attention_pool(sd, p + 'final_attn_pool.', tokens, layer_hiddens, cfg)         # D4:3242-3243
    if final_norm:
        tokens = rmsnorm(tokens, sd[p + 'final_norm.weight'])                                # D4:3247
    return tokens, new_cache


def embed_actions(sd, cfg, prev_actions):
    """ActionEmbedder.forward discrete path + learned embed, D4:1513-1562, 7101.
    prev_actions (b na) int64 -> (b D)."""
    offsets = torch.tensor([0, *torch.tensor(cfg.num_discrete_actions).cumsum(0)[:-1].tolist()])
    emb = sd['action_embedder.discrete_action_embed.weight'][prev_actions + offsets]
    return sd['action_learned_embed'] + emb.sum(dim=-2)


def forward_step(sd, cfg: OracleConfig, noised_latent, signal_level, step_size_log2, prev_actions,
                 kv_cache, token_count, tasks=None):
    """DynamicsWorldModel.forward inference branch restricted to the newest frame,
    D4:6792-7295 / get_prediction D4:7161-7281.  noised_latent (b N Dl).
    Returns (pred latent (b N Dl), agent embed (b D), new kv)."""
    b = noised_latent.shape[0]
    D = cfg.dim
    # latents -> spatial tokens, D4:7168 / 4819-4828
    if cfg.same_len:
        space = noised_latent @ sd['latents_to_spatial_tokens.weight'].T + sd['latents_to_spatial_tokens.bias']
    else:
        space = learned_queries_pool(sd, 'latents_to_spatial_tokens.', noised_latent, cfg.attn_dim_head)
    # flow token D4:7193-7199
    flow = torch.cat((sd['signal_levels_embed.weight'][signal_level], sd['step_size_embed.weight'][step_size_log2]))
    flow = flow[None, None].expand(b, 1, D)
    registers = sd['register_tokens'][None].expand(b, -1, -1)
    agent = sd['agent_learned_embed'][None].expand(b, -1, -1)                                # D4:7004
    if tasks is not None:
        agent = agent + sd['task_embed.weight'][tasks][:, None]                             # D4:7009-7010
    parts = [flow, space, registers]
    if cfg.has_actions:
        if prev_actions is None:
            action = torch.zeros(b, 1, D)                                                    # D4:7110-7115, 7124-7126
        else:
            action = embed_actions(sd, cfg, prev_actions)[:, None]
        parts.append(action)
    parts.append(agent)
    tokens = torch.cat(parts, dim=1)                                                         # D4:7222
    tokens, new_kv = transformer_step(sd, cfg, tokens, kv_cache, token_count)
    space_out = tokens[:, 1:1 + cfg.num_spatial_tokens]
    agent_out = tokens[:, -1]
    # to_latent_pred, D4:4830-4834, 7251
    x = rmsnorm(space_out, sd['to_latent_pred.0.weight'])
    if cfg.same_len:
        pred = x @ sd['to_latent_pred.2.weight'].T
    else:
        x = learned_queries_pool(sd, 'to_latent_pred.1.', x, cfg.attn_dim_head)
        pred = x @ sd['to_latent_pred.2.weight'].T
    return pred, agent_out, new_kv


# --------------------------------------------------------------------------------------
# noise


class TorchRNGNoise:
    """Draws from torch's global generator in the reference's per-frame order and shapes
    (D4:6475, 6611, discrete sampler, 6670) so that a seeded CPU run is draw-for-draw comparable
    with the reference."""

    def latent(self, frame, shape):
        return torch.randn(shape)

    def terminal(self, frame, probs):
        return torch.bernoulli(probs) == 1.0

    def action_uniform(self, frame, type_index, shape):
        return torch.rand(shape)

    def context(self, frame, shape):
        return torch.randn(shape)

    def decoder(self, shape):
        return torch.randn(shape)          # VideoTokenizer.decode's start noise, D4:4204 (drawn after the whole rollout)


class InjectedNoise:
    """Pre-generated noise shared verbatim between the oracle and the CUDA path.
    latent (H b N Dl) normal, action_u (H b A_total) uniform, terminal_u (H b) uniform."""

    def __init__(self, latent, action_u=None, terminal_u=None, decoder_noise=None):
        self.lat, self.act, self.term, self.dec = latent, action_u, terminal_u, decoder_noise
        self._splits = None

    def decoder(self, shape):
        return self.dec.reshape(shape)

    def latent(self, frame, shape):
        return self.lat[frame].reshape(shape).clone()

    def terminal(self, frame, probs):
        return self.term[frame].reshape(probs.shape) < probs

    def action_uniform(self, frame, type_index, shape):
        sizes = self._splits
        start = sum(sizes[:type_index])
        return self.act[frame][..., start:start + sizes[type_index]].reshape(shape)

    def context(self, frame, shape):
        return None       # dead numerically when the time cache is on


# --------------------------------------------------------------------------------------
# generate


@dataclass
class OracleExperience:
    latents: torch.Tensor
    agent_embed: Optional[torch.Tensor] = None
    rewards: Optional[torch.Tensor] = None
    values: Optional[torch.Tensor] = None
    actions: Optional[torch.Tensor] = None                 # discrete (b t na) int64
    log_probs: Optional[torch.Tensor] = None               # discrete (b t na)
    old_action_unembeds: Optional[torch.Tensor] = None     # (b t A_total)
    lens: Optional[torch.Tensor] = None
    is_truncated: Optional[torch.Tensor] = None
    terminals: Optional[torch.Tensor] = None
    step_size: int = 16
    episode_return: Optional[torch.Tensor] = None
    kv_cache: list = field(default_factory=list)           # per time layer (k, v) each (b*S h T d)
    video: Optional[torch.Tensor] = None                   # (b c t h w) when decoded through a tokenizer


def unembed_logits(sd, policy_embed):
    """ActionEmbedder.unembed with pred_head_index=0, D4:1313-1326: (.. 4D) -> (.. A_total)."""
    return policy_embed @ sd['action_embedder.discrete_action_unembed'][:, 0].T


def _log(t, eps=1e-20):
    return t.clamp(min=eps).log()


def cache_from_reference(next_kv_cache):
    """Reference time_cache.main.next_kv_cache (y 2 b*S h T d), D4:3255-3265 -> the oracle's per-time-layer [(k, v)]."""
    if next_kv_cache is None:
        return None
    return [(layer[0], layer[1]) for layer in next_kv_cache]


@torch.no_grad()
def generate(sd, cfg: OracleConfig, time_steps, batch_size, num_steps=4, noise=None, tasks=None,
             return_terminals=False, discrete_temperature=1.0, prompt_latents=None, prompt_actions=None,
             prompt_rewards=None, kv_cache=None, return_agent_actions=True, tokenizer=None, prompt=None,
             return_decoded_video=False):
    """DynamicsWorldModel.generate, use_time_cache=True, D4:6307-6774: the DreamTrainer flags
    (return_rewards_per_frame, return_agent_actions, return_log_probs_and_values; TR:1422-1428) by default;
    `return_agent_actions=False` is the env wrapper's call (env.py:464-484: rewards per frame, actions supplied).
    Unlike the reference it embeds only the newest frame on each pass; outputs are identical because cached
    passes discard everything but the last frame (D4:2960-2961, 6560).

    Prompted rollouts (D4:6377-6402): `prompt_latents` (b P N Dl), `prompt_actions` (b a na), `prompt_rewards` (b P).
    With `kv_cache` (the cache returned for those P frames) decoding resumes at frame P.  Without it the reference
    runs its uncached multi-frame forward over [prompt, new frame] on every pass of the first new frame; the prompt
    frames sit at signal level max_steps-1 with context noise == themselves (6400: past noise is a clone of the prompt,
    so the lerp at 6497 is the identity) and time attention is causal, so that equals one clean pass per prompt frame
    appending to the cache - restated here as that prefill."""
    noise = noise or TorchRNGNoise()
    if isinstance(noise, InjectedNoise):
        noise._splits = list(cfg.num_discrete_actions)
    B, N, Dl = batch_size, cfg.num_latent_tokens, cfg.dim_latent
    assert math.log2(num_steps).is_integer() and 0 < num_steps <= cfg.max_steps
    step_size = cfg.max_steps // num_steps                                                   # D4:6371
    step_log2 = int(math.log2(step_size))                                                    # D4:6940-6942
    reward_codec = HLGauss(cfg.reward_range, cfg.reward_num_bins, cfg.hl_gauss_sigma_to_bin_ratio, cfg.hl_gauss_eps)
    value_codec = HLGauss(cfg.value_range, cfg.value_num_bins, cfg.hl_gauss_sigma_to_bin_ratio, cfg.hl_gauss_eps)
    should_term = return_terminals and cfg.predict_terminals
    want_heads = return_agent_actions and cfg.has_actions
    na = len(cfg.num_discrete_actions)

    # `tokenizer` = (state_dict, oracle.tokenizer_oracle.TokenizerConfig) of the attached VideoTokenizer: a video `prompt`
    # (b c t h w) is tokenized into prompt_latents (D4:6377-6387) and the finished rollout decoded back (D4:6699-6711)
    if prompt is not None:
        from . import tokenizer_oracle
        assert prompt_latents is None, 'cannot pass in both prompt video and prompt latents'
        prompt_latents = tokenizer_oracle.tokenize(*tokenizer, prompt)

    P = 0 if prompt_latents is None else prompt_latents.shape[1]
    latents = [] if P == 0 else [prompt_latents[:, p].reshape(B, N, Dl) for p in range(P)]  # D4:6393-6396
    rewards = [] if prompt_rewards is None else [prompt_rewards[:, p] for p in range(prompt_rewards.shape[1])]   # D4:6444-6447
    # the action history that conditions frame t is decoded[:, :t] right-padded with index 0 (D4:6519-6522); the
    # forward shifts it by one (7111-7120) and uses a zero token when there is no history at all (7124-7126)
    decoded = torch.empty(B, 0, na, dtype=torch.long) if prompt_actions is None else prompt_actions.clone()      # D4:6420-6428

    def prev_actions_for(t):
        if t == 0 or decoded.shape[1] == 0:
            return None
        return decoded[:, t - 1] if t - 1 < decoded.shape[1] else torch.zeros(B, na, dtype=torch.long)

    agent_embeds, values, log_probs, policy_embeds = [], [], [], []
    terminals = torch.zeros(B, dtype=torch.bool)
    lens = torch.full((B,), time_steps)
    if kv_cache is None:
        for p in range(P):                                                                   # prefill: see the docstring
            _, _, kv_cache = forward_step(sd, cfg, latents[p], cfg.max_steps - 1, step_log2, prev_actions_for(p), kv_cache, p, tasks)
    elif P > 0:
        assert kv_cache[0][0].shape[-2] == P, 'the time cache must cover exactly the prompt frames'
    # a time cache WITHOUT prompt latents (reference tests/test_dreamer.py::test_cache_generate): the call imagines time_steps NEW
    # frames on top of the cached ones - their rotary position is offset by the cached count (D4:3010), nothing else changes
    # (the first new frame has no action history of its own: zero action token, D4:7124-7126)
    offset = kv_cache[0][0].shape[-2] if (P == 0 and kv_cache) else 0

    for frame in range(P, time_steps):
        prev_actions = prev_actions_for(frame)
        x = noise.latent(frame, (B, 1, 1, N, Dl)).reshape(B, N, Dl)                          # D4:6475
        for step in range(num_steps + 1):                                                    # D4:6484-6486
            is_last = step == num_steps
            signal = min(step * step_size, cfg.max_steps - 1)                                # D4:6492
            pred, agent, new_kv = forward_step(sd, cfg, x, signal, step_log2, prev_actions, kv_cache, frame + offset, tasks)
            if is_last:
                kv_cache = new_kv                                                            # D4:6545-6546
                break
            tau = signal / cfg.max_steps                                                     # D4:5413
            x = x + (pred - x) / (1.0 - tau) * (step_size / cfg.max_steps)                   # D4:6567-6580
        # heads on the clean-pass agent token
        reward_logits = rmsnorm(agent, sd['to_reward_pred.nets.0.0.weight']) @ sd['to_reward_pred.nets.0.1.weight'].T
        rewards.append(reward_codec.from_logits(reward_logits))                              # D4:6598-6601
        if should_term:                                                                      # D4:6605-6616
            pooled = x.mean(dim=1)
            logit = mlp(sd, 'to_state_terminal_pred.0.', pooled, cfg.head_activation)[..., 0]
            is_term = noise.terminal(frame, logit.sigmoid())
            just = is_term & ~terminals
            lens = lens.masked_fill(just, frame + 1)
            terminals = terminals | is_term
        agent_embeds.append(agent)
        if want_heads:
            pe = mlp(sd, 'policy_head.', agent, cfg.head_activation)                         # D4:6628
            policy_embeds.append(pe)
            logits = unembed_logits(sd, pe).split(list(cfg.num_discrete_actions), dim=-1)     # D4:1330-1334
            sampled, lps = [], []
            for ti, l in enumerate(logits):
                u = noise.action_uniform(frame, ti, (B, 1, l.shape[-1])).reshape(B, -1)
                g = -_log(-_log(u))                                                          # D4:485-497 form
                idx = ((l / max(discrete_temperature, 1e-10)) + g).argmax(dim=-1)
                sampled.append(idx)
                lps.append(l.log_softmax(dim=-1).gather(-1, idx[:, None])[:, 0])
            decoded = torch.cat((decoded, torch.stack(sampled, dim=-1)[:, None]), dim=1)      # D4:6645
            log_probs.append(torch.stack(lps, dim=-1))
            vb = mlp(sd, 'value_head.', agent, cfg.head_activation)                          # D4:6659-6660
            values.append(value_codec.from_logits(vb))
        latents.append(x)
        noise.context(frame, (B, 1, 1, N, Dl))                                               # D4:6670 (consumes RNG; value dead)
        if should_term and bool(terminals.all()):                                            # D4:6681
            break

    T = len(latents)
    lat = torch.stack(latents, dim=1).clamp(-1.0, 1.0) if T > 0 else torch.empty(B, 0, N, Dl)  # D4:6686 (prompt frames included)
    rew = torch.stack(rewards, dim=1) if rewards else torch.empty(B, 0)
    step_mask = torch.arange(T)[None, :] < lens[:, None]
    exp = OracleExperience(
        latents=lat,
        agent_embed=torch.stack(agent_embeds, dim=1) if agent_embeds else None,              # new frames only (D4:6620-6621)
        rewards=rew,
        lens=lens, is_truncated=~terminals, terminals=terminals, step_size=step_size,
        episode_return=(rew * step_mask.float()).sum(dim=-1) if rew.shape[1] == T else None,  # D4:6741-6743
        kv_cache=kv_cache,
    )
    if return_decoded_video:
        from . import tokenizer_oracle
        tcfg = tokenizer[1]
        start = noise.decoder((B, tcfg.channels, T, tcfg.image_height, tcfg.image_width))
        exp.video = tokenizer_oracle.decode(*tokenizer, lat, noise=start)                    # the clamped latents, D4:6686-6711
    if want_heads:
        exp.actions = decoded                                                                # prompt + sampled (D4:6764)
        if log_probs:
            exp.log_probs = torch.stack(log_probs, dim=1)
            exp.values = torch.stack(values, dim=1)
            exp.old_action_unembeds = unembed_logits(sd, torch.stack(policy_embeds, dim=1))  # D4:6749-6750
    return exp


@torch.no_grad()
def interact_with_env(sd, cfg: OracleConfig, tokenizer, env, num_steps=4, max_timesteps=16, env_is_vectorized=False, seed=None, noise=None):
    """DynamicsWorldModel.interact_with_env (D4:5470-5889) for image observations through the attached VideoTokenizer
    (`tokenizer` = (state_dict, TokenizerConfig)), discrete actions, use_time_cache=True: per env step, tokenize the newest
    frame over the tokenizer's own time cache (5588), one clean pass of the world model over its time cache conditioned on
    the previous action (5612-5628), value (5647-5650), policy -> sample -> log-prob (5652-5673), env.step; when the episode
    is cut by max_timesteps the final observation is evaluated once more for the bootstrap value and every per-step record
    is right-padded by one (5790-5853).  Returns an OracleExperience (lens = episode lengths, is_from_world_model False)."""
    from . import tokenizer_oracle
    tsd, tcfg = tokenizer
    noise = noise or TorchRNGNoise()
    if isinstance(noise, InjectedNoise):
        noise._splits = list(cfg.num_discrete_actions)
    step_size = cfg.max_steps // num_steps
    step_log2 = int(math.log2(step_size))
    value_codec = HLGauss(cfg.value_range, cfg.value_num_bins, cfg.hl_gauss_sigma_to_bin_ratio, cfg.hl_gauss_eps)
    sizes = list(cfg.num_discrete_actions)

    obs = env.reset(seed=seed)
    obs = obs[0] if isinstance(obs, tuple) else obs
    image = lambda o: torch.as_tensor(o, dtype=torch.float32) if env_is_vectorized else torch.as_tensor(o, dtype=torch.float32)[None]
    frame = image(obs)                                                                       # (b c h w)
    B = frame.shape[0]
    latents, agent_embeds, policy_embeds, rewards, values, actions, log_probs = [], [], [], [], [], [], []
    terminated_any = torch.zeros(B, dtype=torch.bool)
    truncated_any = torch.zeros(B, dtype=torch.bool)
    done = torch.zeros(B, dtype=torch.bool)
    lens = torch.zeros(B, dtype=torch.long)
    kv_cache = tok_cache = None
    step = 0

    def observe(frame, t, prev):
        nonlocal kv_cache, tok_cache
        lat, tok_cache = tokenizer_oracle.tokenize_step(tsd, tcfg, frame, tok_cache, t)
        _, agent, kv_cache = forward_step(sd, cfg, lat, cfg.max_steps - 1, step_log2, prev, kv_cache, t, None)
        return lat, agent

    while not bool(done.all()):
        step += 1
        lat, agent = observe(frame, step - 1, actions[-1] if actions else None)
        latents.append(lat)
        values.append(value_codec.from_logits(mlp(sd, 'value_head.', agent, cfg.head_activation)))
        pe = mlp(sd, 'policy_head.', agent, cfg.head_activation)
        policy_embeds.append(pe)
        sampled, lps = [], []
        for ti, l in enumerate(unembed_logits(sd, pe).split(sizes, dim=-1)):
            u = noise.action_uniform(step - 1, ti, (B, 1, l.shape[-1])).reshape(B, -1)
            idx = (l - _log(-_log(u))).argmax(dim=-1)                                        # temperature 1 (5657)
            sampled.append(idx)
            lps.append(l.log_softmax(dim=-1).gather(-1, idx[:, None])[:, 0])
        act = torch.stack(sampled, dim=-1)
        actions.append(act)
        log_probs.append(torch.stack(lps, dim=-1))
        out_action = act.numpy() if env_is_vectorized else act[0].numpy()                    # 5679-5690
        if not env_is_vectorized and out_action.size == 1:
            out_action = int(out_action.item())
        next_obs, reward, terminated, truncated, *_ = env.step(out_action)
        terminated = torch.as_tensor(terminated).reshape(B)
        truncated = torch.as_tensor(truncated).reshape(B)
        lens = torch.where(done, lens, lens + 1)                                             # 5726
        terminated_any |= terminated
        truncated_any |= truncated
        if step >= max_timesteps:
            truncated_any |= ~terminated_any                                                 # 5731-5732
        done |= terminated_any | truncated_any
        rewards.append(torch.as_tensor(reward, dtype=torch.float32).reshape(B))
        agent_embeds.append(agent)
        frame = image(next_obs)
        need_bootstrap = truncated_any & ~terminated_any
        if bool(done.all()) and bool(need_bootstrap.any()):                                  # 5790-5853
            lat, agent = observe(frame, step, actions[-1])
            values.append(value_codec.from_logits(mlp(sd, 'value_head.', agent, cfg.head_activation)))
            rewards.append(torch.zeros(B))
            actions.append(torch.zeros_like(actions[-1]))
            log_probs.append(torch.zeros_like(log_probs[-1]))
            latents.append(lat)
            agent_embeds.append(agent)
            policy_embeds.append(mlp(sd, 'policy_head.', agent, cfg.head_activation))
            lens = torch.where(need_bootstrap, lens + 1, lens)
            break

    rew = torch.stack(rewards, dim=1)
    T = rew.shape[1]
    step_mask = torch.arange(T)[None, :] < lens[:, None]
    return OracleExperience(
        latents=torch.stack(latents, dim=1), agent_embed=torch.stack(agent_embeds, dim=1), rewards=rew,
        values=torch.stack(values, dim=1), actions=torch.stack(actions, dim=1), log_probs=torch.stack(log_probs, dim=1),
        old_action_unembeds=unembed_logits(sd, torch.stack(policy_embeds, dim=1)), lens=lens, is_truncated=truncated_any,
        terminals=terminated_any, step_size=step_size, episode_return=(rew * step_mask.float()).sum(dim=-1), kv_cache=kv_cache)


# --------------------------------------------------------------------------------------
# learn_from_experience


def lens_to_mask(lens, max_len):
    return torch.arange(max_len)[None, :] < lens[:, None]


@torch.no_grad()
def calc_gae(rewards, values, masks, learn_masks, gamma, lam):
    """D4:1566-1600 with the AssocScan recurrence out_t = delta_t + gate_t * out_{t+1} (PARITY UNPINNED)."""
    masks = masks.float()
    values_p = F.pad(values, (0, 1), value=0.0)
    v, v_next = values_p[..., :-1], values_p[..., 1:]
    delta = rewards + gamma * v_next * masks - v
    delta = delta.masked_fill(~learn_masks, 0.0)
    gates = gamma * lam * masks
    gae = torch.zeros_like(delta)
    acc = torch.zeros_like(delta[..., 0])
    for t in reversed(range(delta.shape[-1])):
        acc = delta[..., t] + gates[..., t] * acc
        gae[..., t] = acc
    return gae + v


def masked_mean_all(t, mask):
    """torch-einops-utils masked_mean with dim=None (PARITY UNPINNED): mean over selected elements."""
    return t[mask].mean() if bool(mask.any()) else t[mask].sum()


def learn_from_experience(sd, cfg: OracleConfig, exp: OracleExperience, eps=1e-6, objective='ppo', normalize_advantages=None, ema_stats=None):
    """DynamicsWorldModel.learn_from_experience, objective 'ppo' | 'spo' | 'pmpo', only_learn_policy_value_heads=True,
    stored agent embeds, D4:5893-6305.  `sd` tensors that require grad receive gradients.
    `ema_stats` (keep_reward_ema_stats=True, D4:5987-6013): dict(mean, var: 0-d tensors updated in place, decay=reward_ema_decay,
    quantiles=reward_quantile_filter).  Returns (total_policy_loss, value_loss, aux dict)."""
    assert objective in ('ppo', 'spo', 'pmpo'), f'unknown objective {objective}'             # D4:6214-6215
    B, T = exp.latents.shape[:2]
    rewards, old_values = exp.rewards, exp.values
    mask_for_gae = lens_to_mask(exp.lens, T)                                                 # D4:5946-5950
    rewards = rewards.masked_fill(~mask_for_gae, 0.0)
    old_values = old_values.masked_fill(~mask_for_gae, 0.0)
    learnable_lens = exp.lens - exp.is_truncated.long()                                      # D4:5954-5955
    mask = lens_to_mask(learnable_lens, T)
    gae_masks = lens_to_mask((exp.lens - 1).clamp(min=0), T)                                 # D4:5959
    if exp.terminals is not None:                                                            # D4:5961-5967
        pos = (exp.lens - 1).clamp(min=0)
        term_seq = (torch.arange(T)[None, :] == pos[:, None]) & exp.terminals[:, None]
        gae_masks = gae_masks.masked_fill(term_seq, False)
    returns = calc_gae(rewards, old_values, gae_masks, mask, cfg.gae_discount_factor, cfg.gae_lambda)
    if ema_stats is not None:                                                                # D4:5987-6013
        decay = 1. - ema_stats['decay']
        rs = returns[mask]
        lo, hi = torch.quantile(rs, torch.tensor(ema_stats['quantiles'], dtype=rs.dtype)).tolist()
        rs = rs.clamp(lo, hi)
        ema_stats['mean'].lerp_(rs.mean(), decay)
        ema_stats['var'].lerp_(rs.var(correction=0), decay)
        std = ema_stats['var'].clamp(min=1e-5).sqrt()
        advantage = (returns - ema_stats['mean']) / std - (old_values - ema_stats['mean']) / std
    else:
        advantage = returns - old_values                                                     # D4:6017
    if normalize_advantages is None:                                                         # D4:6021: pmpo keeps raw advantages
        normalize_advantages = objective != 'pmpo'
    if normalize_advantages:
        mean = masked_mean_all(advantage, mask)                                              # z_score D4:404-410
        var = masked_mean_all((advantage - mean).pow(2), mask)
        advantage = (advantage - mean) / var.clamp(min=eps).sqrt()

    agent = exp.agent_embed.detach()                                                         # D4:6074-6075
    pe = mlp(sd, 'policy_head.', agent, cfg.head_activation)                                 # D4:6080
    logits = unembed_logits(sd, pe).split(list(cfg.num_discrete_actions), dim=-1)
    lps, ents = [], []
    for ti, l in enumerate(logits):                                                          # D4:6090
        lp = l.log_softmax(dim=-1)
        lps.append(lp.gather(-1, exp.actions[..., ti:ti + 1])[..., 0])
        ents.append(-(lp.exp() * lp).sum(dim=-1))
    log_probs = torch.stack(lps, dim=-1).sum(dim=-1)                                         # D4:6111
    entropies = torch.stack(ents, dim=-1)
    old_log_probs = exp.log_probs.sum(dim=-1)                                                # D4:6113-6114
    gate = ((-log_probs * advantage) / cfg.delight_temperature).sigmoid().detach()           # D4:6119-6120
    if objective == 'pmpo':                                                                  # D4:6127-6182
        glp = log_probs * gate if cfg.use_delight_gating else log_probs
        pos = (advantage >= 0.) & mask                                                       # D4:6028-6030, 6137-6142
        neg = ~(advantage >= 0.) & mask
        scaled = glp * advantage.tanh().abs()
        num = max(1., mask.sum().item())
        policy_loss = -cfg.pmpo_pos_to_neg_weight * (scaled[pos].sum() - scaled[neg].sum()) / num
        if cfg.pmpo_kl_div_loss_weight > 0.:                                                 # D4:6158-6182
            # the reference hands kl_div the FLAT (b t A_total) unembeds (D4:6160-6169, no return_split_discrete), so
            # MultiCategorical sees one categorical over the concatenated action types - restated as called
            new_flat = torch.cat(logits, dim=-1).log_softmax(dim=-1)
            old_flat = exp.old_action_unembeds.log_softmax(dim=-1)
            src, tgt = (old_flat, new_flat) if cfg.pmpo_reverse_kl else (new_flat, old_flat)  # kl_div(src, tgt) D4:1463-1485
            kl = (src.exp() * (src - tgt)).sum(dim=-1)
            policy_loss = policy_loss + masked_mean_all(kl, mask) * cfg.pmpo_kl_div_loss_weight
    else:
        ratio = (log_probs - old_log_probs).exp()
        if objective == 'spo':                                                               # D4:6184-6198
            policy_loss = -(ratio * advantage - advantage.abs() * (ratio - 1.).square() / (2 * cfg.ppo_eps_clip))
        else:                                                                                # D4:6200-6212
            clipped = ratio.clamp(1.0 - cfg.ppo_eps_clip, 1.0 + cfg.ppo_eps_clip)
            policy_loss = -torch.min(ratio * advantage, clipped * advantage)
        if cfg.use_delight_gating:
            policy_loss = policy_loss * gate
        policy_loss = masked_mean_all(policy_loss, mask)
    entropy_loss = masked_mean_all(-entropies.sum(dim=-1), mask)                             # D4:6219-6221
    total_policy_loss = policy_loss + entropy_loss * cfg.policy_entropy_weight               # D4:6238-6242

    value_codec = HLGauss(cfg.value_range, cfg.value_num_bins, cfg.hl_gauss_sigma_to_bin_ratio, cfg.hl_gauss_eps)
    value_bins = mlp(sd, 'value_head.', agent, cfg.head_activation)                          # D4:6268
    return_bins = value_codec.to_probs(returns)                                              # D4:6277
    value_loss = -(return_bins * value_bins.log_softmax(dim=-1)).sum(dim=-1)                 # D4:6281
    value_loss = value_loss[mask].mean()                                                     # D4:6295
    aux = dict(returns=returns, advantage=advantage, log_probs=log_probs, entropies=entropies, mask=mask)
    return total_policy_loss, value_loss, aux


def config_from_reference_kwargs(**kw) -> OracleConfig:
    """Maps DynamicsWorldModel.__init__ kwargs (D4:4662-4778) onto OracleConfig."""
    attn_kwargs = kw.get('attn_kwargs', {}) or {}
    rk = kw.get('reward_encoder_kwargs', {}) or {}
    vk = kw.get('value_encoder_kwargs', None)
    vk = rk if vk is None else vk
    ffk = kw.get('ff_kwargs', {}) or {}
    return OracleConfig(
        dim=kw['dim'], dim_latent=kw['dim_latent'], num_latent_tokens=kw['num_latent_tokens'],
        depth=kw.get('depth', 4), num_spatial_tokens=kw.get('num_spatial_tokens', 4),
        num_register_tokens=kw.get('num_register_tokens', 8), time_block_every=kw.get('time_block_every', 4),
        attn_heads=kw.get('attn_heads', 8), attn_dim_head=kw.get('attn_dim_head', 64),
        query_heads=attn_kwargs.get('query_heads'), attn_softclamp_value=kw.get('attn_softclamp_value', 50.0),
        max_steps=kw.get('max_steps', 64), num_discrete_actions=kw.get('num_discrete_actions', 0),
        multi_token_pred_len=kw.get('multi_token_pred_len', 8),
        policy_head_mlp_depth=kw.get('policy_head_mlp_depth', 3), value_head_mlp_depth=kw.get('value_head_mlp_depth', 3),
        ff_activation=ffk.get('activation', 'silu'), ff_expansion_factor=ffk.get('expansion_factor', 4.0),
        reward_range=rk.get('reward_range', (-20.0, 20.0)), reward_num_bins=rk.get('num_bins', 255),
        value_range=vk.get('reward_range', (-20.0, 20.0)), value_num_bins=vk.get('num_bins', 255),
        predict_terminals=kw.get('predict_terminals', True), num_tasks=kw.get('num_tasks', 0),
        gae_discount_factor=kw.get('gae_discount_factor', 0.997), gae_lambda=kw.get('gae_lambda', 0.95),
        ppo_eps_clip=kw.get('ppo_eps_clip', 0.2), policy_entropy_weight=kw.get('policy_entropy_weight', 0.01),
        use_delight_gating=kw.get('use_delight_gating', True), delight_temperature=kw.get('delight_temperature', 1.0),
        pmpo_pos_to_neg_weight=kw.get('pmpo_pos_to_neg_weight', 0.5), pmpo_reverse_kl=kw.get('pmpo_reverse_kl', True),
        pmpo_kl_div_loss_weight=kw.get('pmpo_kl_div_loss_weight', 0.3),
    )
