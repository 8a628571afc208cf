"""Derives the engine's packed weights from a reference-layout state_dict (host-side plumbing, runs once per
weight change — never on the per-frame path).

Packed names (what d4_bind() resolves; shapes in the reference's nn.Linear (out, in) layout):

  sig_emb, step_emb, registers, agent_embed, action_learned, action_emb, task_emb, inv_freq
  l2s.{w_kv, q, gate, k_gamma, w_out}      latents -> spatial tokens, learned-query pool (reference dreamer4.py:2179-2210, 4819-4828)
  l2s.{w, b}                               ... or the plain Linear when num_spatial_tokens == num_latent_tokens
  vr.w                                     to_value_residual (3026-3027), RMSNorm gamma folded into the weight
  L{i}.attn.{w, b, k_gamma, w_out}         fused rows [to_q; to_k; to_v; to_gates; value-residual mix], gamma folded (1968-2075)
  L{i}.ff.{w_in, b_in, w_out, b_out}       GLU rows interleaved [x0, g0, x1, g1, ...], gamma folded; w_out zero-padded to ff_inner_pad (2105-2116)
  P{i}.{w_qg, w_kv, k_gamma, w_out}, PF.*  attention-residual pools (2143-2177); norm / norm_context gammas folded
  FA.{w_qg, w_kv, k_gamma, w_out}, FAFF.*  final agent cross-attention + feed-forward (3227-3238)
  lp.{norm0, norm_ctx, w_kv, q, gate, k_gamma, w_comb} | lp.{norm0, w}     to_latent_pred (4830-4834); w_comb = Linear(D->Dl) @ to_out
  reward.{w, centers}, value.centers       reward Ensemble member 0 (5067-5075), HL-Gauss bin centres (1041-1105)
  policy.{l}.{w,b,lnw,lnb}, value.{l}.*, terminal.{l}.*, unembed            borrowed parameter pointers (no copy: optimizer steps stay visible)

Every GEMM weight `name` also gets `name.hi` / `name.lo` (tf32 split: hi has 10 explicit mantissa bits, lo = w - hi) when
the engine runs in tf32x3 precision, and on top of those `name.h16hi` / `name.h16lo` + a scale (f16_split) in the experimental
f16x3 precision.

Folding gamma:  rmsnorm(x; gamma) @ W^T == rstd(x) * (x @ (W * gamma)^T): the engine computes rstd per row and applies it as
the GEMM's row scale, so normalised activations never round-trip through HBM."""
from __future__ import annotations

import torch
import torch.nn.functional as F

GEMM_WEIGHTS_SUFFIXES = ('.w', '.w_kv', '.w_out', '.w_in', '.w_qg', '.w_comb')


def _rms(x, w):
    return F.rms_norm(x, (x.shape[-1],), w, None)


def tf32_round(w):
    """Nearest TF32 value (10 explicit mantissa bits, ties away from zero — what `cvt.rna.tf32.f32` does), as fp32."""
    return ((w.contiguous().view(torch.int32) + 0x1000) & -8192).view(torch.float32)


def tf32_split(w):
    """w = hi + lo with hi = rna_tf32(w), lo = rna_tf32(w - hi): round-to-nearest keeps the dropped lo*lo term of the
    3xTF32 product unbiased (a truncating split leaves same-signed errors that add up coherently over K)."""
    hi = tf32_round(w)
    return hi, tf32_round(w - hi)


def f16_split(w):
    """fp16 operand split of the experimental f16x3 mode (csrc/gemm_f16.cu): q = the power of two that brings rms(q w) to ~1,
    hi = fp16(q w), lo = fp16(q w - hi) - both round-to-nearest, so hi + lo carries ~22 significand bits of q w wherever lo stays
    a normal fp16 number.  Returns (hi, lo, 1 / q); the GEMM's epilogue multiplies by 1 / q (an exact exponent shift)."""
    rms = w.float().pow(2).mean().sqrt().clamp_min(1e-30)
    q = torch.exp2(torch.round(torch.log2(1.0 / rms)))
    wq = w.float() * q
    hi = wq.half()
    lo = (wq - hi.float()).half()
    return hi.contiguous(), lo.contiguous(), float(1.0 / q)


def hl_gauss_tables(lo, hi, num_bins, device):
    support = torch.linspace(lo, hi, num_bins + 1).float()
    centers = (support[:-1] + support[1:]) / 2
    return support.to(device), centers.to(device)


def _attn_qg(g, p, queries=None):
    wq, wg, nw = g(p + 'to_q.weight'), g(p + 'to_gates.0.weight'), g(p + 'norm.weight')
    if queries is not None:          # learned queries: the query side is input independent -> precompute
        xq = _rms(queries, nw)
        return xq @ wq.T, xq @ wg.T
    return torch.cat((wq, wg)) * nw[None, :]


def _pack_transformer(g, out, tp, depth, ff_inner_pad, device, attn_qg=None):
    """Packed weights of one AxialSpaceTimeTransformer whose state_dict keys start with `tp` (the dynamics model's
    'transformer.', the video tokenizer's 'encoder_transformer.' / 'decoder.transformer.'): vr.w, L{i}.attn.*, L{i}.ff.*,
    P{i}.*, PF.*, FA.*, FAFF.* (names in the module docstring)."""
    attn_qg = attn_qg or (lambda p, queries=None: _attn_qg(g, p, queries))
    out['vr.w'] = g(tp + 'to_value_residual.1.weight') * g(tp + 'to_value_residual.0.weight')[None, :]

    def pack_ff(p, name):
        nw = g(p + 'norm.weight')
        w_in, b_in = g(p + 'proj_in.weight'), g(p + 'proj_in.bias')
        inner = w_in.shape[0] // 2
        xw, gw = w_in[:inner], w_in[inner:]
        out[name + '.w_in'] = torch.stack((xw, gw), dim=1).reshape(2 * inner, -1) * nw[None, :]
        out[name + '.b_in'] = torch.stack((b_in[:inner], b_in[inner:]), dim=1).reshape(-1)
        out[name + '.w_out'] = F.pad(g(p + 'proj_out.weight'), (0, ff_inner_pad - inner))
        out[name + '.b_out'] = g(p + 'proj_out.bias')

    def pack_pool(p, name):
        out[name + '.w_qg'] = attn_qg(p)
        out[name + '.w_kv'] = torch.cat((g(p + 'to_k.weight'), g(p + 'to_v.weight'))) * g(p + 'norm_context.weight')[None, :]
        out[name + '.k_gamma'] = g(p + 'k_heads_rmsnorm.gamma')
        out[name + '.w_out'] = g(p + 'to_out.weight')

    for i in range(depth):
        p = f'{tp}layers.{i}.2.fn.'
        nw = g(p + 'norm.weight')
        mixw, mixb = g(p + 'to_learned_value_residual_mix.0.weight'), g(p + 'to_learned_value_residual_mix.0.bias')
        w = torch.cat((g(p + 'to_q.weight'), g(p + 'to_k.weight'), g(p + 'to_v.weight'), g(p + 'to_gates.0.weight'), mixw))
        out[f'L{i}.attn.w'] = w * nw[None, :]
        out[f'L{i}.attn.b'] = torch.cat((torch.zeros(w.shape[0] - mixb.shape[0], device=device), mixb))
        out[f'L{i}.attn.k_gamma'] = g(p + 'k_heads_rmsnorm.gamma')
        out[f'L{i}.attn.w_out'] = g(p + 'to_out.weight')
        pack_ff(f'{tp}layers.{i}.3.fn.', f'L{i}.ff')
        if i != depth - 1:
            pack_pool(f'{tp}attn_pools.{i}.fn.attn.', f'P{i}')
    pack_pool(tp + 'final_attn_pool.fn.attn.', 'PF')
    pack_pool(tp + 'final_special_cross_attn.fn.', 'FA')
    pack_ff(tp + 'final_special_ff.fn.', 'FAFF')


def pack(sd, cfg, device, agent_index=0, split=False, split_f16=False):
    """sd: reference-layout state_dict (tensors on `device`); cfg: dreamer4_b200.dynamics.ModelConfig.
    Returns {packed name: fp32 contiguous tensor}; with `split_f16` also `name.h16hi` / `name.h16lo` (fp16) per GEMM weight and
    their 1 / q under the key 'h16scales' (a {name: float} dict, not a tensor)."""
    g = lambda k: sd[k].detach().to(device=device, dtype=torch.float32)
    out = {}
    D, Dl = cfg.dim, cfg.dim_latent
    out['sig_emb'] = g('signal_levels_embed.weight')
    out['step_emb'] = g('step_size_embed.weight')
    out['registers'] = g('register_tokens')
    out['agent_embed'] = g('agent_learned_embed')[agent_index]
    if cfg.has_actions:
        out['action_learned'] = g('action_learned_embed')[agent_index]
        out['action_emb'] = g('action_embedder.discrete_action_embed.weight')
    if cfg.num_tasks > 0:
        out['task_emb'] = g('task_embed.weight')
    out['inv_freq'] = g('transformer.time_rotary.inv_freq')

    attn_qg = lambda p, queries=None: _attn_qg(g, p, queries)

    if cfg.same_len:
        out['l2s.w'] = g('latents_to_spatial_tokens.weight')
        out['l2s.b'] = g('latents_to_spatial_tokens.bias')
        out['lp.w'] = g('to_latent_pred.2.weight')
    else:
        p = 'latents_to_spatial_tokens.attn.'
        out['l2s.q'], out['l2s.gate'] = attn_qg(p, g('latents_to_spatial_tokens.queries'))
        out['l2s.w_kv'] = torch.cat((g(p + 'to_k.weight'), g(p + 'to_v.weight'))) * g(p + 'norm_context.weight')[None, :]
        out['l2s.k_gamma'] = g(p + 'k_heads_rmsnorm.gamma')
        out['l2s.w_out'] = g(p + 'to_out.weight')
        p = 'to_latent_pred.1.attn.'
        out['lp.q'], out['lp.gate'] = attn_qg(p, g('to_latent_pred.1.queries'))
        out['lp.norm_ctx'] = g(p + 'norm_context.weight')
        out['lp.w_kv'] = torch.cat((g(p + 'to_k.weight'), g(p + 'to_v.weight')))
        out['lp.k_gamma'] = g(p + 'k_heads_rmsnorm.gamma')
        out['lp.w_comb'] = g('to_latent_pred.2.weight') @ g(p + 'to_out.weight')
    out['lp.norm0'] = g('to_latent_pred.0.weight')
    _pack_transformer(g, out, 'transformer.', cfg.depth, cfg.ff_inner_pad, device, attn_qg)

    out['reward.w'] = g('to_reward_pred.nets.0.1.weight') * g('to_reward_pred.nets.0.0.weight')[None, :]
    _, out['reward.centers'] = hl_gauss_tables(*cfg.reward_range, cfg.reward_num_bins, device)
    _, out['value.centers'] = hl_gauss_tables(*cfg.value_range, cfg.value_num_bins, device)

    out = {k: v.contiguous() for k, v in out.items()}
    if split:
        for k in list(out):
            if k.endswith(GEMM_WEIGHTS_SUFFIXES) and not k.startswith('reward.'):
                out[k + '.hi'], out[k + '.lo'] = (t.contiguous() for t in tf32_split(out[k]))
    if split_f16:
        scales = {}
        for k in list(out):
            if k.endswith(GEMM_WEIGHTS_SUFFIXES) and not k.startswith('reward.'):
                out[k + '.h16hi'], out[k + '.h16lo'], scales[k] = f16_split(out[k])
        out['h16scales'] = scales
    return out


def mlp_param_names(prefix, n_layers):
    """(packed name, state_dict key) pairs of an x-mlps normed MLP whose pointers are borrowed as-is."""
    pairs = []
    for l in range(n_layers):
        pairs += [(f'{l}.w', f'{prefix}.layers.{l}.0.weight'), (f'{l}.b', f'{prefix}.layers.{l}.0.bias')]
        if l < n_layers - 1:
            pairs += [(f'{l}.lnw', f'{prefix}.layers.{l}.1.weight'), (f'{l}.lnb', f'{prefix}.layers.{l}.1.bias')]
    return pairs


# ------------------------------------------------------------------------------------------------ video tokenizer

def _split_all(out, split, split_f16=False):
    """tf32 hi / lo words per GEMM weight (`split`); with `split_f16` also the fp16 words of the f16x3 mode and their 1 / q under
    the key 'h16scales' (a {name: float} dict, not a tensor), as pack() does."""
    out = {k: v.contiguous() for k, v in out.items()}
    names = [k for k in out if k.endswith(GEMM_WEIGHTS_SUFFIXES)]
    if split:
        for k in names:
            out[k + '.hi'], out[k + '.lo'] = (t.contiguous() for t in tf32_split(out[k]))
    if split_f16:
        scales = {}
        for k in names:
            out[k + '.h16hi'], out[k + '.h16lo'], scales[k] = f16_split(out[k])
        out['h16scales'] = scales
    return out


def _mlp_eval(g, p, x, act):
    """x-mlps normed MLP (Linear -> LayerNorm -> act)* -> Linear on the host: only for input-independent tables at pack time."""
    fn = F.silu if act == 'silu' else F.gelu
    i = 0
    while True:
        try:
            w, b = g(p + f'layers.{i}.0.weight'), g(p + f'layers.{i}.0.bias')
        except KeyError:
            return x
        x = x @ w.T + b
        try:
            x = fn(F.layer_norm(x, (x.shape[-1],), g(p + f'layers.{i}.1.weight'), g(p + f'layers.{i}.1.bias')))
        except KeyError:
            pass
        i += 1


def pack_tokenizer(sd, cfg, device, split=False, split_f16=False):
    """Packed weights of the VideoTokenizer's inference paths (reference dreamer4.py:3686-4237, default branches), from a
    reference-layout state_dict; cfg: dreamer4_b200.tokenizer.TokenizerConfig.  Returns {'enc': {...}, 'dec': {...}, 'io': {...}}:

      enc / dec   one transformer each (_pack_transformer names + inv_freq + final_norm), what d4_tf_bind resolves
      io          patch.{w, b, ln}       patch_to_tokens: Linear(p*p*c -> D) + LayerNorm(no bias) (3833-3838)
                  latent_tokens (N, D)   the encoder's special tokens (4349)
                  to_latents.w (Dl, D)   encoded_to_latents, followed by tanh (4413, 4426)
                  npatch.{w, b, ln}      noised_patch_to_tokens of the flow decoder (3881-3886)
                  pos_emb (hp*wp, D)     to_decoder_pos_emb(coords): input independent -> evaluated here (3618-3623)
                  lat_in.w (D, Dl)       latents_to_decoder (4148-4154); time_embed (steps, D) is its per-flow-step bias
                  to_patch.{w, b}        decoder.tokens_to_patch (3569-3572)"""
    g = lambda k: sd[k].detach().to(device=device, dtype=torch.float32)
    enc, dec, io = {}, {}, {}
    for out, tp, depth in ((enc, 'encoder_transformer.', cfg.encoder_depth), (dec, 'decoder.transformer.', cfg.decoder_depth)):
        _pack_transformer(g, out, tp, depth, cfg.ff_inner_pad, device)
        out['inv_freq'] = g(tp + 'time_rotary.inv_freq')
        out['final_norm'] = g(tp + 'final_norm.weight')
    io['patch.w'], io['patch.b'], io['patch.ln'] = g('patch_to_tokens.1.weight'), g('patch_to_tokens.1.bias'), g('patch_to_tokens.2.weight')
    io['latent_tokens'] = g('latent_tokens')
    io['to_latents.w'] = g('encoded_to_latents.weight')
    io['npatch.w'], io['npatch.b'], io['npatch.ln'] = (g('noised_patch_to_tokens.1.weight'), g('noised_patch_to_tokens.1.bias'),
                                                       g('noised_patch_to_tokens.2.weight'))
    hp, wp = cfg.image_height // cfg.patch_size, cfg.image_width // cfg.patch_size
    coords = torch.stack(torch.meshgrid(torch.linspace(-1., 1., hp), torch.linspace(-1., 1., wp), indexing='ij'), dim=-1).to(device)
    io['pos_emb'] = _mlp_eval(g, 'decoder.to_decoder_pos_emb.', coords.reshape(hp * wp, 2), cfg.decoder_pos_emb_mlp_activation)
    io['lat_in.w'] = g('latents_to_decoder.weight')
    io['time_embed'] = g('time_embed.weight')
    io['to_patch.w'], io['to_patch.b'] = g('decoder.tokens_to_patch.0.weight'), g('decoder.tokens_to_patch.0.bias')
    return dict(enc=_split_all(enc, split, split_f16), dec=_split_all(dec, split, split_f16), io=_split_all(io, split))
