"""CPU restatement of the reference VideoTokenizer's inference paths - TEST INFRASTRUCTURE, like dreamer4_oracle.py.

The steps either side of the rollout: `tokenize` (video -> latents, feeds generate(prompt=...),
D4:4107-4113, 4239-4431) and `decode` (latents -> video, generate(return_decoded_video=True), D4:4137-4237, 3490-3682),
default branches only (no causal conv3d, shifted patches, slot attention, aug conditioning, separate flow decoder, MoT, RNN,
H-Net, PoPE).  PARITY: pinned by tests/golden/tokenizer/*.pt, produced by the reference's own source (oracle/make_golden.py);
the composition with the rollout (generate(prompt=video), return_decoded_video) is pinned through dreamer4_oracle.generate.
There is no CUDA path for the tokenizer yet - this is the checker, built first.

Both transformers are AxialSpaceTimeTransformers with causal time attention, so - exactly like the dynamics model's rollout -
a T-frame forward equals T single-frame steps over a growing time-KV cache; the restatement runs them that way through
dreamer4_oracle.transformer_step, which is also the shape a B200 decode kernel path would take (one pass per frame)."""
from dataclasses import dataclass

import torch
import torch.nn.functional as F

from . import dreamer4_oracle as O


@dataclass
class TokenizerConfig:
    """VideoTokenizer.__init__ kwargs on the default path (D4:3686-3766)."""
    dim: int
    dim_latent: int
    patch_size: int
    image_height: int
    image_width: int
    num_latent_tokens: int = 64
    encoder_depth: int = 4
    decoder_depth: int = 4
    time_block_every: int = 4
    attn_heads: int = 8
    attn_dim_head: int = 64
    attn_softclamp_value: float = 50.0
    ff_activation: str = 'silu'
    channels: int = 3
    decoder_flow_steps: int = 1
    decoder_pos_emb_mlp_activation: str = 'silu'
    pool_dim_head: int = 64

    @property
    def encoder(self):                 # D4:3908-3929
        return _TransformerCfg(self.encoder_depth, self.time_block_every, self.attn_heads, self.attn_dim_head, self.attn_softclamp_value,
                               self.ff_activation, self.pool_dim_head)

    @property
    def decoder(self):                 # D4:3595-3607: VideoDecoderNetwork does not forward ff_kwargs / attn_kwargs / softclamp -> library defaults
        return _TransformerCfg(self.decoder_depth, self.time_block_every, self.attn_heads, self.attn_dim_head, 50.0, 'silu', self.pool_dim_head)


@dataclass
class _TransformerCfg:
    depth: int
    time_block_every: int
    attn_heads: int
    attn_dim_head: int
    attn_softclamp_value: float
    ff_activation: str
    pool_dim_head: int

    @property
    def is_time(self):                 # D4:2845
        return [((i + 1) % self.time_block_every) == 0 for i in range(self.depth)]


def config_from_reference_kwargs(**kw) -> TokenizerConfig:
    size = kw.get('image_size')
    ffk = kw.get('ff_kwargs', {}) or {}
    return TokenizerConfig(
        dim=kw['dim'], dim_latent=kw['dim_latent'], patch_size=kw['patch_size'],
        image_height=kw.get('image_height') or size, image_width=kw.get('image_width') or size,
        num_latent_tokens=kw.get('num_latent_tokens', 64), encoder_depth=kw.get('encoder_depth', 4),
        decoder_depth=kw.get('decoder_depth', 4), time_block_every=kw.get('time_block_every', 4),
        attn_heads=kw.get('attn_heads', 8), attn_dim_head=kw.get('attn_dim_head', 64),
        attn_softclamp_value=kw.get('attn_softclamp_value', 50.0), ff_activation=ffk.get('activation', 'silu'),
        channels=kw.get('channels', 3), decoder_flow_steps=kw.get('decoder_flow_steps', 1),
        decoder_pos_emb_mlp_activation=kw.get('decoder_pos_emb_mlp_activation', 'silu'))


def patchify(frame, p):
    """'b c (h p1) (w p2) -> b (h w) (p1 p2 c)' (D4:3835, 3883)."""
    b, c, H, W = frame.shape
    x = frame.reshape(b, c, H // p, p, W // p, p)
    return x.permute(0, 2, 4, 3, 5, 1).reshape(b, (H // p) * (W // p), p * p * c)


def unpatchify(patches, p, c, H, W):
    """'b (h w) (p1 p2 c) -> b c (h p1) (w p2)' (D4:3571)."""
    b = patches.shape[0]
    x = patches.reshape(b, H // p, W // p, p, p, c)
    return x.permute(0, 5, 1, 3, 2, 4).reshape(b, c, H, W)


def patch_tokens(sd, prefix, frame, p):
    """Sequential(Rearrange, Linear, LayerNorm(bias=False)) (D4:3833-3838, 3881-3886)."""
    x = patchify(frame, p) @ sd[prefix + '1.weight'].T + sd[prefix + '1.bias']
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + '2.weight'], None)


@torch.no_grad()
def tokenize_step(sd, cfg: TokenizerConfig, frame, cache, t):
    """One frame (b c h w) through the encoder over its time-KV cache `cache` (None at t = 0) -> (latents (b n dl), cache):
    the incremental call interact_with_env makes per env step (D4:5588, forward(..., time_cache=, return_time_cache=True))."""
    N = cfg.num_latent_tokens
    latent_tokens = sd['latent_tokens'][None].expand(frame.shape[0], -1, -1)                 # D4:4349
    tokens = torch.cat((patch_tokens(sd, 'patch_to_tokens.', frame, cfg.patch_size), latent_tokens), dim=1)                   # D4:4360
    # encoder: the N latent tokens are the special tokens - patches cannot attend to them, they attend to everything,
    # and cross-attend to the patches once more at the end (D4:3912-3918, 1769-1783, 3227-3238)
    tokens, cache = O.transformer_step(sd, cfg.encoder, tokens, cache, t, prefix='encoder_transformer.', num_special=N, final_norm=True)
    return (tokens[:, -N:] @ sd['encoded_to_latents.weight'].T).tanh(), cache               # D4:4413, 4426


@torch.no_grad()
def tokenize(sd, cfg: TokenizerConfig, video):
    """VideoTokenizer.tokenize = forward(video, return_latents=True) in eval mode (no patch masking): (b c t h w) -> (b t n dl)."""
    if video.ndim == 4:                                                                      # D4:4258-4260
        video = video[:, :, None]
    cache, out = None, []
    for t in range(video.shape[2]):
        latents, cache = tokenize_step(sd, cfg, video[:, :, t], cache, t)
        out.append(latents)
    return torch.stack(out, dim=1)


@torch.no_grad()
def decode(sd, cfg: TokenizerConfig, latents, noise=None):
    """VideoTokenizer.decode (D4:4183-4237) with the default single flow step: (b t n dl) -> (b c t h w).
    `noise` (b c t h w) stands in for the randn at D4:4204 (drawn from torch's global generator when None)."""
    assert cfg.decoder_flow_steps >= 1, 'the plain (non-flow) decoder is decoder_flow_steps=0'
    b, T = latents.shape[:2]
    c, p, H, W = cfg.channels, cfg.patch_size, cfg.image_height, cfg.image_width
    noise = torch.randn(b, c, T, H, W) if noise is None else noise
    steps = cfg.decoder_flow_steps
    times = torch.linspace(0., 1., steps + 1)
    tcfg = cfg.decoder
    hp, wp = H // p, W // p
    coords = torch.stack(torch.meshgrid(torch.linspace(-1., 1., hp), torch.linspace(-1., 1., wp), indexing='ij'), dim=-1)
    pos_emb = O.mlp(sd, 'decoder.to_decoder_pos_emb.', coords.reshape(hp * wp, 2), cfg.decoder_pos_emb_mlp_activation)        # D4:3618-3623
    video = noise
    for i in range(steps):
        latent_tokens = latents @ sd['latents_to_decoder.weight'].T + sd['time_embed.weight'][i]                              # D4:4148-4154
        cache, frames = None, []
        for t in range(T):
            spatial = pos_emb[None] + patch_tokens(sd, 'noised_patch_to_tokens.', video[:, :, t], p)                          # D4:3625-3628
            tokens = torch.cat((spatial, latent_tokens[:, t]), dim=1)                                                         # D4:3655
            # decoder transformer: library defaults - ONE special token (the last latent), final norm (D4:3595-3607, 2774)
            tokens, cache = O.transformer_step(sd, tcfg, tokens, cache, t, prefix='decoder.transformer.', num_special=1, final_norm=True)
            patches = tokens[:, :hp * wp] @ sd['decoder.tokens_to_patch.0.weight'].T + sd['decoder.tokens_to_patch.0.bias']  # D4:3569-3572
            frames.append(unpatchify(patches, p, c, H, W))
        pred = torch.stack(frames, dim=2)
        flow = (pred - video) / (1. - times[i])                                                                               # D4:4223-4227
        video = video + flow * (1. / steps)
    return video
