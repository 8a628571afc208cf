"""Host mirror of the reference `VideoTokenizer`'s INFERENCE paths (reference dreamer4/dreamer4.py:3686-4237): `tokenize`
(video -> latents, what `generate(prompt=...)` and `interact_with_env` call) and `decode` (latents -> video, what
`generate(return_decoded_video=True)` calls), default branches only.  Same constructor keywords and `state_dict` keys as the
reference class, so a reference checkpoint loads with `strict=True`.

Like the dynamics model, both of its transformers are run ONE FRAME PER STEP over a time-KV cache (causal time attention makes
that equal to the reference's multi-frame forward; oracle/tokenizer_oracle.py holds this against the reference's golden
vectors), through the native library: `d4_tf_step` (a generic AxialSpaceTimeTransformer frame step with `num_special` special
tokens and a final norm) between a patch-embedding front end and a projection back end made of `d4_patchify`,
`d4_linear_rows`, `d4_tok_assemble`, `d4_tanh_rows` and `d4_unpatchify_flow` (include/d4b200.h).  There is no CPU fallback.

STATUS (round 1): written after the round's GPU budget was spent - the CUDA side compiles for sm_100a but has never run on
hardware.  On the CPU, this class + engine.cu's d4_tf_step + the new kernels' device code reproduce the reference's golden
tokenize / decode vectors under the CUDA-thread simulator (tests/test_kernels_cusim_cpu.py), and the packed dataflow does so
in torch (tests/test_tokenizer_cpu.py); the GPU parity tests (tests/test_zy_tokenizer_gpu.py) are non-strict xfail until
their first hardware run."""
from __future__ import annotations

import ctypes as C
import functools
import pickle
from collections import namedtuple
from dataclasses import dataclass

import torch
from torch import nn

from . import _lib
from ._lib import D4Error, check, ptr
from .dynamics import _linear_b, _linear_w
from .packing import pack_tokenizer


@dataclass
class TokenizerConfig:
    """VideoTokenizer.__init__ keywords on the default path (reference dreamer4.py:3686-3766)."""
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
    decoder_pos_mlp_depth: int = 2
    pool_heads: int = 4
    pool_dim_head: int = 64

    @property
    def num_patches(self):
        return (self.image_height // self.patch_size) * (self.image_width // self.patch_size)

    @property
    def dim_patch(self):
        return self.patch_size * self.patch_size * self.channels

    @property
    def tokens_per_frame(self):            # patches + latent tokens, encoder (4360) and decoder (3655) alike
        return self.num_patches + self.num_latent_tokens

    @property
    def ff_inner(self):                    # reference dreamer4.py:2094
        return int(self.dim * 4 * 2 / 3)

    @property
    def ff_inner_pad(self):
        return (self.ff_inner + 31) // 32 * 32

    def is_time(self, depth):              # reference dreamer4.py:2845
        return [((i + 1) % self.time_block_every) == 0 for i in range(depth)]


# keywords of the reference constructor whose non-default values select branches this path does not build
_UNSUPPORTED = dict(
    use_causal_conv3d=False, use_shifted_patches=False, use_slot_attention=False, has_aug_conditioning=False,
    separate_flow_decoder=False, mot_temporal=False, use_time_rnn=False, time_attention_use_pope=False, num_video_views=1,
)
# accepted and ignored: they shape the training losses only
_TRAINING_ONLY = ('lpips_loss_weight', 'lpips_loss_network', 'encoder_add_decor_aux_loss', 'decor_auxx_loss_weight', 'per_image_patch_mask_prob',
                  'latent_ar_loss_weight', 'recon_loss_weight', 'norm_recon_loss', 'time_decorr_loss_weight', 'nd_rotary_kwargs')


# what `forward(..., return_time_cache=True)` hands back and `time_cache=` takes: the encoder's keys / values stay in the engine's
# in-place buffer, the record only says how many frames it holds and which rollout wrote them (a later un-cached call restarts
# the buffer and makes older records stale - rejected rather than silently mixing rollouts)
TokenizerTimeCache = namedtuple('TokenizerTimeCache', ['token_count', 'epoch'])


def _register(lib, ctx, packed):
    """Hands a packed weight dict to a native context: tensors by name, and - f16x3 mode - the 1 / q of each fp16-split weight."""
    for name, t in packed.items():
        if name != 'h16scales':
            check(lib.d4_set_weight(ctx, name.encode(), ptr(t), t.numel()))
    for name, scale in packed.get('h16scales', {}).items():
        check(lib.d4_set_weight_scale(ctx, name.encode(), scale))


def _records_config(init):
    @functools.wraps(init)
    def wrapped(self, *args, **kwargs):
        self._config = (args, kwargs)
        init(self, *args, **kwargs)
    return wrapped


class VideoTokenizer(nn.Module):
    """Drop-in for the reference class on `tokenize` / `decode`.  Extra keyword (not in the reference): `precision` in
    {'tf32x3' (default), 'f16x3', 'fp32', 'tf32'} as for DynamicsWorldModel: arithmetic of the transformers' linear layers ('f16x3': the
    fp16 3-term split GEMM, 30.5 k vs 25.8 k frames/s tokenize at 128 videos; the patch / latent projections stay on 3xTF32); attention inside
    a frame runs on 3xTF32 tensor-core tiles in every mode but 'fp32' (csrc/frame_attn_mma.cu)."""

    @_records_config
    def __init__(self, dim, dim_latent, patch_size, image_height=None, image_width=None, image_size=None, num_latent_tokens=64,
                 encoder_depth=4, decoder_depth=4, time_block_every=4, attn_kwargs: dict = dict(), attn_dim_head=64, attn_heads=8,
                 attn_softclamp_value=50., ff_kwargs: dict = dict(), channels=3, decoder_flow_steps=1,
                 decoder_pos_emb_mlp_activation='silu', decoder_pos_mlp_depth=2, precision='tf32x3', time_attn_variant=1, **kwargs):
        super().__init__()
        for k, dflt in _UNSUPPORTED.items():
            if kwargs.pop(k, dflt) != dflt:
                raise NotImplementedError(f'VideoTokenizer({k}=...) is outside the path this package builds (DESIGN.md section 8)')
        for k in _TRAINING_ONLY:
            kwargs.pop(k, None)
        if kwargs:
            raise TypeError(f'unknown / unsupported VideoTokenizer keywords: {sorted(kwargs)}')
        if attn_kwargs:
            raise NotImplementedError('VideoTokenizer(attn_kwargs=...) is outside the path this package builds')
        image_height, image_width = image_height or image_size, image_width or image_size
        assert image_height and image_width, 'image_size or image_height / image_width is required'
        assert image_height % patch_size == 0 and image_width % patch_size == 0
        assert decoder_flow_steps >= 1, 'the plain (non-flow) decoder is outside the path this package builds'
        assert precision in ('fp32', 'tf32', 'tf32x3', 'f16x3')
        self.cfg = TokenizerConfig(dim=dim, dim_latent=dim_latent, patch_size=patch_size, image_height=image_height, image_width=image_width,
                                   num_latent_tokens=num_latent_tokens, encoder_depth=encoder_depth, decoder_depth=decoder_depth,
                                   time_block_every=time_block_every, attn_heads=attn_heads, attn_dim_head=attn_dim_head,
                                   attn_softclamp_value=attn_softclamp_value, ff_activation=(ff_kwargs or {}).get('activation', 'silu'),
                                   channels=channels, decoder_flow_steps=decoder_flow_steps,
                                   decoder_pos_emb_mlp_activation=decoder_pos_emb_mlp_activation, decoder_pos_mlp_depth=decoder_pos_mlp_depth)
        self.precision, self.time_attn_variant = precision, time_attn_variant
        self.image_height, self.image_width, self.channels, self.patch_size = image_height, image_width, channels, patch_size
        self.num_latent_tokens, self.dim_latent, self.decoder_flow_steps = num_latent_tokens, dim_latent, decoder_flow_steps
        self._build_parameters()
        self._ctx = {}            # 'enc' / 'dec' -> (ctx, key, buffers)
        self._packed, self._packed_version = None, None
        self._enc_epoch = 0

    # ------------------------------------------------------------------ parameters (reference state_dict layout)

    _reg = None                   # filled in below from DynamicsWorldModel's helpers

    def _build_parameters(self):
        c = self.cfg
        D, Dl, h, d, P = c.dim, c.dim_latent, c.attn_heads, c.attn_dim_head, c.dim_patch
        self._reg('latent_tokens', torch.randn(c.num_latent_tokens, D) * 1e-2)
        self._reg('mask_token', torch.randn(D) * 1e-2)
        for name in ('patch_to_tokens', 'noised_patch_to_tokens'):
            self._reg(f'{name}.1.weight', _linear_w(D, P))
            self._reg(f'{name}.1.bias', _linear_b(D, P))
            self._reg(f'{name}.2.weight', torch.ones(D))
        self._reg('time_embed.weight', torch.randn(c.decoder_flow_steps, D))
        self._reg_transformer('encoder_transformer.', c.encoder_depth)
        self._reg('encoded_to_latents.weight', _linear_w(Dl, D))
        self._reg('latents_to_decoder.weight', _linear_w(D, Dl))
        self._reg_mlp('decoder.to_decoder_pos_emb', (2, *((2 * D,) * (c.decoder_pos_mlp_depth + 1)), D))      # create_mlp(depth) (3526-3532)
        self._reg('decoder.tokens_to_patch.0.weight', _linear_w(P, D))
        self._reg('decoder.tokens_to_patch.0.bias', _linear_b(P, D))
        self._reg_transformer('decoder.transformer.', c.decoder_depth)
        self._reg('recon_loss_normalizer.exp_avg_sq', torch.ones(1), buffer=True)

    def _reg_transformer(self, p, depth):
        c = self.cfg
        D, h, d = c.dim, c.attn_heads, c.attn_dim_head
        inv_freq = 1.0 / (10000. ** (torch.arange(0, d, 2).float() / d))
        self._reg(p + 'time_rotary.inv_freq', inv_freq, buffer=True)
        self._reg(p + 'to_value_residual.0.weight', torch.ones(D))
        self._reg(p + 'to_value_residual.1.weight', _linear_w(h * d, D))
        for i in range(depth):
            self._reg_attention(f'{p}layers.{i}.2.fn.', D, D, h, h, d, False, True)
            self._reg_ff(f'{p}layers.{i}.3.fn.', D, c.ff_inner)
        for i in range(depth - 1):
            self._reg_attention(f'{p}attn_pools.{i}.fn.attn.', D, D, c.pool_heads, c.pool_heads, c.pool_dim_head, True, False)
        self._reg_attention(p + 'final_attn_pool.fn.attn.', D, D, c.pool_heads, c.pool_heads, c.pool_dim_head, True, False)
        self._reg(p + 'final_norm.weight', torch.ones(D))
        self._reg_attention(p + 'final_special_cross_attn.fn.', D, D, h, h, d, True, True)
        self._reg_ff(p + 'final_special_ff.fn.', D, c.ff_inner)

    @property
    def device(self):
        return self.latent_tokens.device

    # .save / .load / .init_and_load of the reference's @save_load (dreamer4.py:3684)

    def save(self, path, overwrite=True):
        import os
        assert overwrite or not os.path.exists(str(path)), f'{path} already exists'
        torch.save(dict(model=self.state_dict(), config=pickle.dumps(self._config)), str(path))

    def load(self, path, strict=True):
        self.load_state_dict(torch.load(str(path), map_location='cpu', weights_only=False)['model'], strict=strict)

    @classmethod
    def init_and_load(cls, path, strict=True):
        pkg = torch.load(str(path), map_location='cpu', weights_only=False)
        args, kwargs = pickle.loads(pkg['config'])
        model = cls(*args, **kwargs)
        model.load_state_dict(pkg['model'], strict=strict)
        return model

    # ------------------------------------------------------------------ engine plumbing

    def _version(self):
        return sum(p._version for p in self.parameters())

    def _release(self):
        lib = _lib.load() if self._ctx else None
        for ctx, _, _ in self._ctx.values():
            lib.d4_ctx_destroy(ctx)
        self._ctx, self._packed, self._packed_version = {}, None, None

    def __deepcopy__(self, memo):
        """A copy (DynamicsWorldModel(copy_video_tokenizer=True), reference dreamer4.py:4789-4792) owns no native context."""
        import copy
        fresh = dict(_ctx={}, _packed=None, _packed_version=None)
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        new.__dict__ = {k: (fresh[k] if k in fresh else copy.deepcopy(v, memo)) for k, v in self.__dict__.items()}
        return new

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._release()
        return out

    def _require_cuda(self):
        if self.device.type != 'cuda':
            raise D4Error('dreamer4_b200 runs on CUDA only: move the tokenizer to a B200 (`.cuda()`); there is no CPU fallback')

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _weights(self):
        if self._packed_version != self._version():
            if self._ctx:
                self._release()          # contexts hold pointers into the old packed tensors
            self._packed = pack_tokenizer(self.state_dict(), self.cfg, self.device, split=self.precision in ('tf32x3', 'f16x3'),
                                          split_f16=self.precision == 'f16x3')          # f16x3: odd shapes and the io GEMMs stay on 3xTF32
            self._packed_version = self._version()
        return self._packed

    def _transformer(self, which, batch, need_time, keep_frames=0):
        """Native context of the encoder ('enc') or decoder ('dec') transformer for this batch with a KV capacity of at least
        `need_time` frames (allocated in blocks of 16); when an existing context has to grow, its first `keep_frames` cached
        frames are carried over."""
        self._require_cuda()
        lib = _lib.load()
        c, dev = self.cfg, self.device
        packed = self._weights()
        key = (batch, self.precision, self.time_attn_variant, dev.index)
        have = self._ctx.get(which)
        if have is not None and have[1] == key and have[2]['max_time'] >= need_time:
            return lib, have[0], have[2]
        old = have if (have is not None and have[1] == key and keep_frames > 0) else None
        max_time = (need_time + 15) // 16 * 16
        depth = c.encoder_depth if which == 'enc' else c.decoder_depth
        cc = _lib.d4_tf_config()
        cc.dim, cc.depth, cc.time_block_every = c.dim, depth, c.time_block_every
        cc.heads, cc.query_heads, cc.dim_head = c.attn_heads, c.attn_heads, c.attn_dim_head
        cc.pool_heads, cc.pool_dim_head = c.pool_heads, c.pool_dim_head
        # VideoDecoderNetwork does not forward ff_kwargs / softclamp to its transformer (reference dreamer4.py:3595-3607): defaults
        act = c.ff_activation if which == 'enc' else 'silu'
        cc.ff_inner, cc.ff_inner_pad, cc.ff_act = c.ff_inner, c.ff_inner_pad, 1 if act == 'gelu' else 0
        cc.tokens_per_frame = c.tokens_per_frame
        cc.num_special = c.num_latent_tokens if which == 'enc' else 1          # 3912-3918 / library default
        cc.final_norm = 1
        cc.softclamp = c.attn_softclamp_value if which == 'enc' else 50.0
        cc.max_batch, cc.max_time = batch, max_time
        cc.precision, cc.time_attn_variant = _lib.PREC[self.precision], self.time_attn_variant
        ctx = C.c_void_p()
        check(lib.d4_tf_create(C.byref(cc), C.byref(ctx)))
        ws_bytes, kv_bytes = lib.d4_workspace_bytes(ctx), lib.d4_kv_bytes(ctx)
        y = max(sum(c.is_time(depth)), 1)
        ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=dev)
        ws = ws[(-ws.data_ptr()) % 256:][:ws_bytes]                      # the engine wants its workspace 256-byte aligned
        kv = torch.zeros(y, 2, batch * c.tokens_per_frame, c.attn_heads, max_time, c.attn_dim_head, device=dev)
        assert kv.numel() * 4 == kv_bytes
        if old is not None:
            kv[..., :keep_frames, :] = old[2]['kv'][..., :keep_frames, :]
        check(lib.d4_set_buffers(ctx, ptr(ws), ws_bytes, ptr(kv), kv_bytes))
        _register(lib, ctx, packed[which])
        check(lib.d4_bind(ctx))
        if have is not None:
            lib.d4_ctx_destroy(have[0])
        bufs = dict(ws=ws, kv=kv, max_time=max_time)
        self._ctx[which] = (ctx, key, bufs)
        return lib, ctx, bufs

    def _linear(self, lib, A, w_name, bias, M, amap=(0, 0, 0), out=None):
        """out (M, N) = A[rows through amap] @ io[w_name]^T + bias on the engine's GEMM (precision of the class)."""
        io = self._packed['io']
        W = io[w_name]
        N, K = W.shape
        out = torch.empty(M, N, device=A.device) if out is None else out
        lo = io.get(w_name + '.lo')
        Wp = io[w_name + '.hi'] if lo is not None else W
        stream = self._stream()
        check(lib.d4_linear_rows(_lib.PREC['tf32x3' if self.precision == 'f16x3' else self.precision], M, N, K, ptr(A), A.stride(-2), amap[0], amap[1], amap[2], ptr(Wp), K, ptr(lo),
                                 ptr(W), ptr(bias), ptr(out), N, stream))
        return out

    # ------------------------------------------------------------------ tokenize (video -> latents)

    @torch.no_grad()
    def tokenize(self, video, time_cache=None, return_time_cache=False):
        """(b c t h w) or (b c h w) -> (b t n dl): `forward(video, return_latents=True)` in eval mode (reference dreamer4.py:4107-4113).
        `time_cache` / `return_time_cache` (as on the reference's forward, used by interact_with_env at :5588): continue over the
        frames already encoded instead of starting a new video."""
        c = self.cfg
        if video.ndim == 4:                                              # reference dreamer4.py:4258-4260
            video = video[:, :, None]
        b, ch, T, H, W = video.shape
        assert (ch, H, W) == (c.channels, c.image_height, c.image_width), f'video {tuple(video.shape)} does not match the tokenizer'
        video = video.to(device=self.device, dtype=torch.float32).contiguous()
        t0 = 0
        if time_cache is not None:
            if time_cache.epoch != self._enc_epoch:
                raise ValueError('stale tokenizer time_cache: a later un-cached tokenize() restarted the encoder\'s in-place KV buffer')
            t0 = time_cache.token_count
        else:
            self._enc_epoch += 1
        lib, ctx, _ = self._transformer('enc', b, t0 + T, keep_frames=t0)
        io = self._packed['io']
        P, N, S, D = c.num_patches, c.num_latent_tokens, c.tokens_per_frame, c.dim
        stream = self._stream()
        patches = torch.empty(b * P, c.dim_patch, device=self.device)
        lin = torch.empty(b * P, D, device=self.device)
        tok, out = torch.empty(b, S, D, device=self.device), torch.empty(b, S, D, device=self.device)
        latents = torch.empty(b, T, N, c.dim_latent, device=self.device)
        for t in range(T):
            frame = video[:, :, t]
            check(lib.d4_patchify(b, ch, H, W, c.patch_size, ptr(frame), frame.stride(0), frame.stride(1), ptr(patches), stream))
            self._linear(lib, patches, 'patch.w', io['patch.b'], b * P, out=lin)
            check(lib.d4_tok_assemble(b, S, P, D, ptr(lin), ptr(io['patch.ln']), None, ptr(io['latent_tokens']), 0, N, ptr(tok), stream))
            check(lib.d4_tf_step(ctx, b, ptr(tok), t0 + t, ptr(out), stream))
            lat_t = torch.empty(b * N, c.dim_latent, device=self.device)
            self._linear(lib, out.view(b * S, D), 'to_latents.w', None, b * N, amap=(N, S, P), out=lat_t)
            check(lib.d4_tanh_rows(ptr(lat_t), lat_t.numel(), stream))
            latents[:, t] = lat_t.view(b, N, c.dim_latent)
        if return_time_cache:
            return latents, TokenizerTimeCache(t0 + T, self._enc_epoch)
        return latents

    # ------------------------------------------------------------------ decode (latents -> video)

    @torch.no_grad()
    def decode(self, latents, height=None, width=None, aug_id=None, return_recons_across_steps=False, noise=None):
        """(b t n dl) -> (b c t h w): the flow decoder's Euler steps from noise (reference dreamer4.py:4183-4237).
        `noise` (b c t h w) replaces the randn drawn at :4204 (extra keyword, for seeded parity tests)."""
        c = self.cfg
        assert aug_id is None and not return_recons_across_steps, 'outside the path this package builds'
        H, W = height or c.image_height, width or c.image_width
        assert (H, W) == (c.image_height, c.image_width), 'decoding at another resolution is outside the path this package builds'
        latents = latents.to(device=self.device, dtype=torch.float32).contiguous()
        b, T, N, Dl = latents.shape
        assert (N, Dl) == (c.num_latent_tokens, c.dim_latent)
        lib, ctx, _ = self._transformer('dec', b, T)
        io = self._packed['io']
        P, S, D, ch = c.num_patches, c.tokens_per_frame, c.dim, c.channels
        stream = self._stream()
        video = (torch.randn(b, ch, T, H, W, device=self.device) if noise is None else noise.to(device=self.device, dtype=torch.float32)).clone()
        steps = c.decoder_flow_steps
        patches = torch.empty(b * P, c.dim_patch, device=self.device)
        lin = torch.empty(b * P, D, device=self.device)
        spec = torch.empty(b * N, D, device=self.device)
        tok, out = torch.empty(b, S, D, device=self.device), torch.empty(b, S, D, device=self.device)
        pred = torch.empty(b * P, c.dim_patch, device=self.device)
        for i in range(steps):
            tau = i / steps                                              # linspace(0, 1, steps + 1)[i]
            scale = (1.0 / (1.0 - tau)) * (1.0 / steps)
            for t in range(T):                                           # every flow step restarts the decoder's time cache
                frame = video[:, :, t]
                check(lib.d4_patchify(b, ch, H, W, c.patch_size, ptr(frame), frame.stride(0), frame.stride(1), ptr(patches), stream))
                self._linear(lib, patches, 'npatch.w', io['npatch.b'], b * P, out=lin)
                lat_t = latents[:, t].reshape(b * N, Dl)
                self._linear(lib, lat_t, 'lat_in.w', io['time_embed'][i], b * N, out=spec)
                check(lib.d4_tok_assemble(b, S, P, D, ptr(lin), ptr(io['npatch.ln']), ptr(io['pos_emb']), ptr(spec), N * D, N, ptr(tok), stream))
                check(lib.d4_tf_step(ctx, b, ptr(tok), t, ptr(out), stream))
                self._linear(lib, out.view(b * S, D), 'to_patch.w', io['to_patch.b'], b * P, amap=(P, S, 0), out=pred)
                check(lib.d4_unpatchify_flow(b, ch, H, W, c.patch_size, ptr(pred), ptr(frame), frame.stride(0), frame.stride(1), scale, stream))
        return video

    def forward(self, *args, **kwargs):
        if kwargs.pop('return_latents', False):
            assert len(args) == 1, 'forward(video, return_latents=True)'
            kwargs.pop('mask_patches', None)           # eval mode never masks patches
            return self.tokenize(args[0], **kwargs)
        raise NotImplementedError('VideoTokenizer training forward is outside the path this package builds (DESIGN.md section 8)')


class AxialSpaceTimeTransformer(nn.Module):
    """The reference's stand-alone transformer block stack (dreamer4.py:2762-3267, exported by `dreamer4/__init__.py`) on its inference
    path: same constructor keywords (default branches), same `state_dict` keys, `forward(tokens (b t s d), cache=None,
    return_intermediates=False)`.  Frames are run one per `d4_tf_step` over the engine's in-place time-KV cache - equal to the
    reference's multi-frame forward because time attention is causal (golden: tests/golden/cache/axial_transformer.pt); `cache`
    continues from the frames already seen.  No autograd: training through it is outside the path this package builds."""

    def __init__(self, dim, depth, attn_heads=8, attn_dim_head=64, attn_softclamp_value=50., time_block_every=4, attn_kwargs: dict = dict(),
                 ff_kwargs: dict = dict(), num_special_tokens=1, final_norm=True, precision='tf32x3', time_attn_variant=1, **kwargs):
        super().__init__()
        unsupported = dict(special_attend_only_itself=False, full_spatial_attn=False, spatial_modules=None, value_residual=True, rnn_time=False,
                           time_attention_use_pope=False, space_attention_use_pope=False, use_attn_pool=True, mot_temporal=False, h_net_layer=None)
        for k, v in kwargs.items():
            if k in unsupported and v != unsupported[k]:
                raise NotImplementedError(f'AxialSpaceTimeTransformer({k}={v!r}) is outside the path this package builds')
        assert not attn_kwargs and precision in ('fp32', 'tf32', 'tf32x3', 'f16x3')
        self.dim, self.depth, self.num_special_tokens, self.has_final_norm = dim, depth, num_special_tokens, final_norm
        self.precision, self.time_attn_variant = precision, time_attn_variant
        # the tokenizer config doubles as the transformer's (only the fields _reg_transformer and the context need)
        self.cfg = TokenizerConfig(dim=dim, dim_latent=0, patch_size=1, image_height=1, image_width=1, num_latent_tokens=num_special_tokens,
                                   encoder_depth=depth, decoder_depth=depth, time_block_every=time_block_every, attn_heads=attn_heads,
                                   attn_dim_head=attn_dim_head, attn_softclamp_value=attn_softclamp_value,
                                   ff_activation=(ff_kwargs or {}).get('activation', 'silu'))
        VideoTokenizer._reg_transformer(self, '', depth)
        if not final_norm:
            del self._modules['final_norm']
        self._ctx, self._packed, self._packed_version, self._epoch = None, None, None, 0

    _reg, _reg_attention, _reg_ff = None, None, None      # filled in below
    _require_cuda = VideoTokenizer._require_cuda
    _stream = VideoTokenizer._stream

    @property
    def device(self):
        return self.to_value_residual._modules['0'].weight.device

    def _release(self):
        if self._ctx is not None:
            _lib.load().d4_ctx_destroy(self._ctx[0])
        self._ctx, self._packed, self._packed_version = None, None, None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._release()           # parameter storage moved: borrowed pointers are stale
        return out

    def _context(self, batch, tokens_per_frame, need_time, keep_frames):
        from .packing import _pack_transformer, _split_all
        self._require_cuda()
        lib, c, dev = _lib.load(), self.cfg, self.device
        version = sum(p._version for p in self.parameters())
        if self._packed_version != version:
            self._release()
            g = lambda k: self.state_dict()[k].detach().to(device=dev, dtype=torch.float32)
            out = {}
            _pack_transformer(g, out, '', self.depth, c.ff_inner_pad, dev)
            out['inv_freq'] = g('time_rotary.inv_freq')
            if self.has_final_norm:
                out['final_norm'] = g('final_norm.weight')
            self._packed, self._packed_version = _split_all(out, self.precision in ('tf32x3', 'f16x3'), self.precision == 'f16x3'), version
        key = (batch, tokens_per_frame, self.precision, self.time_attn_variant, dev.index)
        have = self._ctx
        if have is not None and have[1] == key and have[2]['max_time'] >= need_time:
            return lib, have[0], have[2]
        old = have if (have is not None and have[1] == key and keep_frames > 0) else None
        max_time = (need_time + 15) // 16 * 16
        cc = _lib.d4_tf_config()
        cc.dim, cc.depth, cc.time_block_every = c.dim, self.depth, c.time_block_every
        cc.heads, cc.query_heads, cc.dim_head = c.attn_heads, c.attn_heads, c.attn_dim_head
        cc.pool_heads, cc.pool_dim_head = c.pool_heads, c.pool_dim_head
        cc.ff_inner, cc.ff_inner_pad, cc.ff_act = c.ff_inner, c.ff_inner_pad, 1 if c.ff_activation == 'gelu' else 0
        cc.tokens_per_frame, cc.num_special, cc.final_norm = tokens_per_frame, self.num_special_tokens, int(self.has_final_norm)
        cc.softclamp = c.attn_softclamp_value
        cc.max_batch, cc.max_time = batch, max_time
        cc.precision, cc.time_attn_variant = _lib.PREC[self.precision], self.time_attn_variant
        ctx = C.c_void_p()
        check(lib.d4_tf_create(C.byref(cc), C.byref(ctx)))
        ws_bytes, kv_bytes = lib.d4_workspace_bytes(ctx), lib.d4_kv_bytes(ctx)
        ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=dev)
        ws = ws[(-ws.data_ptr()) % 256:][:ws_bytes]
        y = max(sum(c.is_time(self.depth)), 1)
        kv = torch.zeros(y, 2, batch * tokens_per_frame, c.attn_heads, max_time, c.attn_dim_head, device=dev)
        assert kv.numel() * 4 == kv_bytes
        if old is not None:
            kv[..., :keep_frames, :] = old[2]['kv'][..., :keep_frames, :]
        check(lib.d4_set_buffers(ctx, ptr(ws), ws_bytes, ptr(kv), kv_bytes))
        _register(lib, ctx, self._packed)
        check(lib.d4_bind(ctx))
        if have is not None:
            lib.d4_ctx_destroy(have[0])
        bufs = dict(ws=ws, kv=kv, max_time=max_time)
        self._ctx = (ctx, key, bufs)
        return lib, ctx, bufs

    @torch.no_grad()
    def forward(self, tokens, time_lens=None, cache=None, return_intermediates=False, **kwargs):
        """tokens (b t s d) -> (b t s d) [, TransformerIntermediates(next_kv_cache (y 2 b*s h T d), token_count)].  With `cache`, only
        the LAST frame of `tokens` is new (the reference excises the past ones and hands them back untouched, :2957-2961, 3249-3250)."""
        from .experience import TransformerIntermediates
        assert time_lens is None and tokens.ndim == 4, 'tokens (b t s d)'
        b, T, S, D = tokens.shape
        tokens = tokens.to(device=self.device, dtype=torch.float32)
        t0, new = 0, tokens
        if cache is not None:
            if getattr(cache.next_kv_cache, '_d4_epoch', self._epoch) != self._epoch:
                raise ValueError('stale cache: a later un-cached forward restarted the in-place KV buffer')
            t0 = cache.token_count
            new = tokens[:, -1:] if T > 1 else tokens
        else:
            self._epoch += 1
        lib, ctx, bufs = self._context(b, S, t0 + new.shape[1], keep_frames=t0)
        stream = self._stream()
        out = torch.empty_like(new)
        for t in range(new.shape[1]):
            frame_in, frame_out = new[:, t].contiguous(), torch.empty(b, S, D, device=self.device)
            check(lib.d4_tf_step(ctx, b, ptr(frame_in), t0 + t, ptr(frame_out), stream))
            out[:, t] = frame_out
        if cache is not None and T > 1:
            out = torch.cat((tokens[:, :-1], out), dim=1)
        if not return_intermediates:
            return out
        count = t0 + new.shape[1]
        y = sum(self.cfg.is_time(self.depth))
        kv = bufs['kv'][:y, :, :, :, :count] if y > 0 else None
        if kv is not None:
            kv._d4_epoch = self._epoch
        return out, TransformerIntermediates(next_kv_cache=kv, token_count=count)


# the reference-layout parameter registration helpers are those of the dynamics model
from .dynamics import DynamicsWorldModel as _D  # noqa: E402

VideoTokenizer._reg = _D._reg
VideoTokenizer._reg_attention = _D._reg_attention
VideoTokenizer._reg_ff = _D._reg_ff
VideoTokenizer._reg_mlp = _D._reg_mlp
AxialSpaceTimeTransformer._reg = _D._reg
AxialSpaceTimeTransformer._reg_attention = _D._reg_attention
AxialSpaceTimeTransformer._reg_ff = _D._reg_ff
