"""Torch (CPU) emulation of the native engine's per-pass dataflow (dreamer4_b200/csrc/engine.cu::run_pass) over the
PACKED weights of dreamer4_b200/packing.py.  Test infrastructure: it lets the CPU suite check the packing algebra
(gamma folding, GLU interleave, fused qkv rows, hidden-stack layout, row maps) against the oracle without a GPU.
It shares the buffer names and step order with engine.cu on purpose."""
import torch
import torch.nn.functional as F

EPS = torch.finfo(torch.float32).eps


def rstd(x):
    return torch.rsqrt(x.pow(2).mean(dim=-1) + EPS)


def l2n(x):
    return x / x.norm(dim=-1, keepdim=True).clamp(min=1e-12)


def headnorm(k, gamma):            # k (..., h, d), gamma (h, d)
    return l2n(k) * ((gamma + 1.) * k.shape[-1] ** 0.5)


def small_attn(q, k, v, k_gamma, scale, gate=None, softclamp=0., num_special=0, belief=False, v0=None, mix=None):
    """q (b, nq, hq, d); k, v (b, n, h, d); gate (b, nq, hq) logits; mix (b, n, h) logits."""
    b, nq, hq, d = q.shape
    h = k.shape[2]
    g = hq // h
    if v0 is not None:
        v = torch.lerp(v, v0, torch.sigmoid(mix)[..., None])
    k = headnorm(k, k_gamma)
    kk = k.repeat_interleave(g, dim=2)
    vv = v.repeat_interleave(g, dim=2)
    sim = torch.einsum('bihd,bjhd->bhij', q, kk) * scale
    if softclamp > 0:
        sim = torch.tanh(sim / softclamp) * softclamp
    if num_special:                      # non-special queries do not see the special keys (the last num_special tokens)
        n = k.shape[1]
        m = torch.ones(nq, n, dtype=torch.bool)
        m[:nq - num_special, n - num_special:] = False
        sim = sim.masked_fill(~m, -torch.finfo(sim.dtype).max)
    out = torch.einsum('bhij,bjhd->bihd', sim.softmax(dim=-1), vv)
    if belief:
        vh = l2n(vv)
        out = out - (out * vh).sum(dim=-1, keepdim=True) * vh
    if gate is not None:
        out = out * torch.sigmoid(gate)[..., None]
    return out, k, v


def rope(x, t, inv_freq):            # x (..., d)
    ang = t * inv_freq
    ang = torch.cat((ang, ang))
    x1, x2 = x.chunk(2, dim=-1)
    return x * ang.cos() + torch.cat((-x2, x1), dim=-1) * ang.sin()


def glu(x, act):
    xs, gs = x[..., 0::2], x[..., 1::2]
    return xs * (F.silu(gs) if act == 'silu' else F.gelu(gs))


def emulate_transformer(P, tok, kv_cache, t, *, heads, query_heads, dim_head, depth, is_time, softclamp, ff_activation, ff_inner_pad,
                        pool_heads, pool_dim_head, num_special=1, final_norm=False):
    """One frame of an AxialSpaceTimeTransformer over packed weights (engine.cu: the layer loop of run_pass / tf_step).
    tok (B, S, D) -> (tokens out (B, S, D), new kv list); the last `num_special` tokens of a frame are the special ones."""
    B, S, D = tok.shape
    h, hq, d, L = heads, query_heads, dim_head, depth
    Dq, Dkv, Dp, hp, dp = hq * d, h * d, pool_heads * pool_dim_head, pool_heads, pool_dim_head
    M = B * S
    ns = num_special
    scale = d ** -0.5
    hid = [tok.reshape(M, D)]
    v0 = (hid[0] @ P['vr.w'].T) * rstd(hid[0])[:, None]
    x_in = hid[0]
    new_kv = []
    ti = 0

    def pool(name, xq, n):
        qg = (xq @ P[name + '.w_qg'].T) * rstd(xq)[:, None]
        stack = torch.cat(hid[:n])                                    # (n*M, D)
        kv = (stack @ P[name + '.w_kv'].T) * rstd(stack)[:, None]
        kv = kv.reshape(n, M, 2 * Dp).transpose(0, 1)                # (M, n, 2Dp)
        k, v = kv[..., :Dp].reshape(M, n, hp, dp), kv[..., Dp:].reshape(M, n, hp, dp)
        o, _, _ = small_attn(qg[:, :Dp].reshape(M, 1, hp, dp), k, v, P[name + '.k_gamma'], dp ** -0.5, gate=qg[:, Dp:].reshape(M, 1, hp))
        return xq + o.reshape(M, Dp) @ P[name + '.w_out'].T

    def ff(name, x):
        mid = glu((x @ P[name + '.w_in'].T) * rstd(x)[:, None] + P[name + '.b_in'], ff_activation)
        mid = F.pad(mid, (0, ff_inner_pad - mid.shape[-1]))
        return x + mid @ P[name + '.w_out'].T + P[name + '.b_out']

    for i in range(L):
        row = (x_in @ P[f'L{i}.attn.w'].T) * rstd(x_in)[:, None] + P[f'L{i}.attn.b']
        q, k, v = row[:, :Dq], row[:, Dq:Dq + Dkv], row[:, Dq + Dkv:Dq + 2 * Dkv]
        gate, mix = row[:, Dq + 2 * Dkv:Dq + 2 * Dkv + hq], row[:, Dq + 2 * Dkv + hq:]
        if is_time[i]:
            qh, kh, vh = q.reshape(M, 1, hq, d), k.reshape(M, 1, h, d), v.reshape(M, 1, h, d)
            vh = torch.lerp(vh, v0.reshape(M, 1, h, d), torch.sigmoid(mix).reshape(M, 1, h, 1))
            kh = rope(headnorm(kh, P[f'L{i}.attn.k_gamma']), float(t), P['inv_freq'])
            qh = rope(qh, float(t), P['inv_freq'])
            kn, vn = kh.transpose(1, 2), vh.transpose(1, 2)           # (M, h, 1, d)
            if kv_cache is not None and t > 0:
                kall, vall = torch.cat((kv_cache[ti][0], kn), dim=2), torch.cat((kv_cache[ti][1], vn), dim=2)
            else:
                kall, vall = kn, vn
            new_kv.append((kall, vall))
            g = hq // h
            kk, vv = kall.repeat_interleave(g, dim=1), vall.repeat_interleave(g, dim=1)
            sim = torch.einsum('mhd,mhjd->mhj', qh[:, 0], kk) * scale
            sim = torch.tanh(sim / softclamp) * softclamp
            o = torch.einsum('mhj,mhjd->mhd', sim.softmax(dim=-1), vv)
            vhat = l2n(vh[:, 0].repeat_interleave(g, dim=1))
            o = o - (o * vhat).sum(dim=-1, keepdim=True) * vhat
            o = (o * torch.sigmoid(gate)[..., None]).reshape(M, Dq)
            ti += 1
        else:
            o, _, _ = small_attn(q.reshape(B, S, hq, d), k.reshape(B, S, h, d), v.reshape(B, S, h, d), P[f'L{i}.attn.k_gamma'], scale,
                                 gate=gate.reshape(B, S, hq), softclamp=softclamp, num_special=ns, belief=True,
                                 v0=v0.reshape(B, S, h, d), mix=mix.reshape(B, S, h))
            o = o.reshape(M, Dq)
        hid.append(x_in + o @ P[f'L{i}.attn.w_out'].T)
        hid.append(ff(f'L{i}.ff', hid[-1]))
        if i != L - 1:
            x_in = pool(f'P{i}', hid[-1], 2 * i + 3)
    xf = hid[2 * L].clone().reshape(B, S, D)
    sp = xf[:, S - ns:].reshape(B * ns, D)                             # the special tokens cross-attend to the others once more
    qg = (sp @ P['FA.w_qg'].T) * rstd(sp)[:, None]
    kv = (hid[2 * L] @ P['FA.w_kv'].T) * rstd(hid[2 * L])[:, None]
    kv = kv.reshape(B, S, 2 * Dkv)[:, :S - ns]
    o, _, _ = small_attn(qg[:, :Dq].reshape(B, ns, hq, d), kv[..., :Dkv].reshape(B, S - ns, h, d), kv[..., Dkv:].reshape(B, S - ns, h, d),
                         P['FA.k_gamma'], scale, gate=qg[:, Dq:].reshape(B, ns, hq))
    sp = sp + o.reshape(B * ns, Dq) @ P['FA.w_out'].T
    sp = ff('FAFF', sp)
    xf[:, S - ns:] = sp.reshape(B, ns, D)
    xf = pool('PF', xf.reshape(M, D), 2 * L + 1).reshape(B, S, D)
    if final_norm:
        xf = xf * rstd(xf)[..., None] * P['final_norm']
    return xf, new_kv


def emulate_pass(P, cfg, latent, signal, step_log2, prev_actions, kv_cache, t):
    """P: packed dict; latent (B, N, Dl); kv_cache: list over time layers of (k, v) each (B*S, h, t, d) or None.
    Returns pred (B, N, Dl), agent (B, D), new kv list."""
    B = latent.shape[0]
    S, D, N, nsp = cfg.tokens_per_frame, cfg.dim, cfg.num_latent_tokens, cfg.num_spatial_tokens
    h, hq, d, L = cfg.attn_heads, cfg.query_heads, cfg.attn_dim_head, cfg.depth
    Dq, Dkv, Dp, hp, dp = hq * d, h * d, cfg.pool_heads * cfg.pool_dim_head, cfg.pool_heads, cfg.pool_dim_head
    M = B * S
    scale = d ** -0.5
    tok = torch.zeros(B, S, D)
    if cfg.same_len:
        tok[:, 1:1 + nsp] = latent @ P['l2s.w'].T + P['l2s.b']
    else:
        x = latent.reshape(B * N, -1)
        kvl = (x @ P['l2s.w_kv'].T) * rstd(x)[:, None]
        k, v = kvl[:, :Dkv].reshape(B, N, h, d), kvl[:, Dkv:].reshape(B, N, h, d)
        q = P['l2s.q'].reshape(1, nsp, hq, d).expand(B, -1, -1, -1)
        o, _, _ = small_attn(q, k, v, P['l2s.k_gamma'], scale, gate=P['l2s.gate'][None].expand(B, -1, -1))
        tok[:, 1:1 + nsp] = o.reshape(B, nsp, Dq) @ P['l2s.w_out'].T
    tok[:, 0] = torch.cat((P['sig_emb'][signal], P['step_emb'][step_log2]))
    tok[:, 1 + nsp:1 + nsp + cfg.num_register_tokens] = P['registers']
    if cfg.has_actions:
        if prev_actions is None:
            tok[:, S - 2] = 0.
        else:
            offs = torch.tensor([0, *torch.tensor(cfg.num_discrete_actions).cumsum(0)[:-1].tolist()])
            tok[:, S - 2] = P['action_learned'] + P['action_emb'][prev_actions + offs].sum(dim=1)
    tok[:, S - 1] = P['agent_embed']
    xf, new_kv = emulate_transformer(P, tok, kv_cache, t, heads=h, query_heads=hq, dim_head=d, depth=L, is_time=cfg.is_time,
                                     softclamp=cfg.attn_softclamp_value, ff_activation=cfg.ff_activation, ff_inner_pad=cfg.ff_inner_pad,
                                     pool_heads=cfg.pool_heads, pool_dim_head=cfg.pool_dim_head, num_special=1)
    agent = xf[:, S - 1]
    sp = xf[:, 1:1 + nsp].reshape(B * nsp, D)
    sp_n = sp * rstd(sp)[:, None] * P['lp.norm0']
    if cfg.same_len:
        pred = (sp_n @ P['lp.w'].T).reshape(B, N, -1)
    else:
        sp_n2 = sp_n * rstd(sp_n)[:, None] * P['lp.norm_ctx']
        kv = sp_n2 @ P['lp.w_kv'].T
        q = P['lp.q'].reshape(1, N, hq, d).expand(B, -1, -1, -1)
        o, _, _ = small_attn(q, kv[:, :Dkv].reshape(B, nsp, h, d), kv[:, Dkv:].reshape(B, nsp, h, d), P['lp.k_gamma'], scale,
                             gate=P['lp.gate'][None].expand(B, -1, -1))
        pred = (o.reshape(B * N, Dq) @ P['lp.w_comb'].T).reshape(B, N, -1)
    return pred, agent, new_kv


# ---------------------------------------------------------------------------------------------------------------------------
# Operand-split emulation: a packed GEMM weight wrapped so that `x @ W.T` inside emulate_pass computes what the tensor-core
# kernels compute - each operand split into two narrow words, three of the four cross products summed (exact products; the
# fp32 accumulation is emulated in fp64, i.e. its rounding is left out: it is common to every mode).  `pow2_rows`: rows of x
# are first scaled by 2^round(log2(rstd(x))) as gemm_f16.cu does in rs_mode 1 (undone exactly afterwards).

def _tf32_rna(x):
    return ((x.contiguous().view(torch.int32) + 0x1000) & -8192).view(torch.float32)


def pow2_near(x):
    return ((x.contiguous().view(torch.int32) + 0x00400000) & 0x7F800000).view(torch.float32)


class SplitWeight:
    def __init__(self, w, mode, pow2_rows=True):
        self.mode, self.pow2_rows, self.shape = mode, pow2_rows, w.shape
        if mode == 'tf32x3':
            self.hi = _tf32_rna(w)
            self.lo = _tf32_rna(w - self.hi)
            self.inv_q = 1.0
        elif mode == 'f16x3':
            from dreamer4_b200.packing import f16_split
            self.hi, self.lo, self.inv_q = f16_split(w)
        else:
            raise ValueError(mode)

    @property
    def T(self):
        return self

    def __rmatmul__(self, x):                      # x (M, K) @ W^T -> (M, N)
        lead = x.shape[:-1]
        x = x.float().reshape(-1, x.shape[-1])
        if self.mode == 'tf32x3':
            a_hi = _tf32_rna(x)
            a_lo = _tf32_rna(x - a_hi)
            p = None
        else:
            p = pow2_near(rstd(x)) if self.pow2_rows else torch.ones(x.shape[0])
            a = x * p[:, None]
            a_hi = a.half()
            a_lo = (a - a_hi.float()).half()
        wh, wl = self.hi.double(), self.lo.double()
        out = a_lo.double() @ wh.T + a_hi.double() @ wl.T + a_hi.double() @ wh.T
        if p is not None:
            out = out / p.double()[:, None] * self.inv_q
        return out.float().reshape(*lead, -1)


def split_packed(P, mode, pow2_rows='engine'):
    """The packed dict with every GEMM weight the engine sends to the tensor cores wrapped in a SplitWeight.  pow2_rows:
    True / False for every weight, or 'engine': only the GEMMs whose rows the engine scales by an RMS statistic (rs_mode 1 of
    the pair kernels - the fused q/k/v, value-residual, feed-forward-in and pool q / kv projections); the projections that
    consume attention outputs or GLU activations see their rows as they are."""
    from dreamer4_b200.packing import GEMM_WEIGHTS_SUFFIXES
    normed = ('.attn.w', '.w_in', '.w_qg', '.w_kv')

    def rows(k):
        if pow2_rows != 'engine':
            return bool(pow2_rows)
        return k == 'vr.w' or (k.endswith(normed) and k != 'lp.w_kv')
    return {k: (SplitWeight(v, mode, rows(k)) if k.endswith(GEMM_WEIGHTS_SUFFIXES) and not k.startswith('reward.') else v) for k, v in P.items()}


# ---------------------------------------------------------------------------------------------------------------------------
# Video tokenizer dataflow over pack_tokenizer()'s weights: the call sequence of dreamer4_b200/tokenizer.py (d4_patchify ->
# d4_linear_rows -> d4_tok_assemble -> d4_tf_step -> d4_linear_rows -> d4_tanh_rows | d4_unpatchify_flow) in torch.

def patchify(frame, p):                     # d4_patchify: (B, C, H, W) -> (B * hp * wp, p * p * C), element order (p1 p2 c)
    b, c, H, W = frame.shape
    x = frame.reshape(b, c, H // p, p, W // p, p)
    return x.permute(0, 2, 4, 3, 5, 1).reshape(b * (H // p) * (W // p), p * p * c)


def unpatchify(patches, b, p, c, H, W):     # inverse layout of d4_unpatchify_flow
    x = patches.reshape(b, H // p, W // p, p, p, c)
    return x.permute(0, 5, 1, 3, 2, 4).reshape(b, c, H, W)


def tok_assemble(lin, ln_w, pos_emb, special, B, P):
    """d4_tok_assemble: patch rows = LayerNorm(lin) * ln_w (+ pos_emb), then the special rows."""
    x = F.layer_norm(lin, (lin.shape[-1],), ln_w, None).reshape(B, P, -1)
    if pos_emb is not None:
        x = x + pos_emb[None]
    special = special[None].expand(B, -1, -1) if special.ndim == 2 else special
    return torch.cat((x, special), dim=1)


def _tf_kwargs(cfg, which):
    depth = cfg.encoder_depth if which == 'enc' else cfg.decoder_depth
    return dict(heads=cfg.attn_heads, query_heads=cfg.attn_heads, dim_head=cfg.attn_dim_head, depth=depth, is_time=cfg.is_time(depth),
                softclamp=cfg.attn_softclamp_value if which == 'enc' else 50.0, ff_activation=cfg.ff_activation if which == 'enc' else 'silu',
                ff_inner_pad=cfg.ff_inner_pad, pool_heads=cfg.pool_heads, pool_dim_head=cfg.pool_dim_head,
                num_special=cfg.num_latent_tokens if which == 'enc' else 1, final_norm=True)


def emulate_tokenize(PK, cfg, video):
    """video (b c t h w) -> latents (b t n dl)."""
    io, enc = PK['io'], PK['enc']
    b, T = video.shape[0], video.shape[2]
    P, N = cfg.num_patches, cfg.num_latent_tokens
    kv, out = None, []
    for t in range(T):
        lin = patchify(video[:, :, t], cfg.patch_size) @ io['patch.w'].T + io['patch.b']
        tok = tok_assemble(lin, io['patch.ln'], None, io['latent_tokens'], b, P)
        x, kv = emulate_transformer(enc, tok, kv, t, **_tf_kwargs(cfg, 'enc'))
        out.append((x[:, P:].reshape(b * N, -1) @ io['to_latents.w'].T).tanh().reshape(b, N, -1))
    return torch.stack(out, dim=1)


def emulate_decode(PK, cfg, latents, noise):
    """latents (b t n dl), noise (b c t h w) -> video (b c t h w)."""
    io, dec = PK['io'], PK['dec']
    b, T, N, _ = latents.shape
    P, p, c, H, W = cfg.num_patches, cfg.patch_size, cfg.channels, cfg.image_height, cfg.image_width
    video = noise.clone()
    steps = cfg.decoder_flow_steps
    for i in range(steps):
        scale = (1.0 / (1.0 - i / steps)) * (1.0 / steps)
        kv = None
        for t in range(T):
            frame = video[:, :, t]
            lin = patchify(frame, p) @ io['npatch.w'].T + io['npatch.b']
            spec = (latents[:, t].reshape(b * N, -1) @ io['lat_in.w'].T + io['time_embed'][i]).reshape(b, N, -1)
            tok = tok_assemble(lin, io['npatch.ln'], io['pos_emb'], spec, b, P)
            x, kv = emulate_transformer(dec, tok, kv, t, **_tf_kwargs(cfg, 'dec'))
            pred = x[:, :P].reshape(b * P, -1) @ io['to_patch.w'].T + io['to_patch.b']
            video[:, :, t] = frame + (unpatchify(pred, b, p, c, H, W) - frame) * scale
    return video
