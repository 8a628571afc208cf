"""The fp16 3-term split GEMM (dreamer4_b200/csrc/gemm_f16.cu) through d4_linear(precision = D4_PREC_F16X3), against fp64, held to
the SAME tolerance as the 3xTF32 kernel in tests/test_gpu_parity.py::test_linear_tcgen05 - the point of the kernel is that accuracy
at half the tensor-core cost (scripts/split_precision_study.py) - and the engine in its f16x3 mode
(DynamicsWorldModel(precision='f16x3')), which routes the transformer's dense layers to that kernel, against the oracle at the tf32x3
mode's tolerances.  First hardware run in round 2 (all green); the full-horizon parity of the mode is in test_horizon_parity_gpu.py."""
import os
import ctypes as C
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

D4_PREC_F16X3 = 3


def pack_f16(W):
    """hi / lo fp16 words of q W (q the power of two that brings rms(q W) to ~1), leading dimension padded to 8, and 1 / q."""
    q = 2.0 ** torch.round(torch.log2(1.0 / W.pow(2).mean().sqrt()))
    Wq = W * q
    N, K = W.shape
    ld = (K + 7) // 8 * 8
    hi = torch.zeros(N, ld, dtype=torch.float16, device=W.device)
    lo = torch.zeros(N, ld, dtype=torch.float16, device=W.device)
    hi[:, :K] = Wq.half()
    lo[:, :K] = (Wq - hi[:, :K].float()).half()
    return hi, lo, ld, float(1.0 / q)


@pytest.mark.parametrize('M,N,K,act', [(256, 256, 64, 0), (128, 128, 512, 0), (300, 200, 96, 0), (1000, 1552, 512, 0), (515, 2730, 512, 1),
                                      (700, 512, 1376, 0), (4096, 260, 512, 0), (130, 170, 64, 2), (3840, 1024, 32, 0), (2048, 2048, 2048, 0)])
def test_linear_f16x3(M, N, K, act):
    from dreamer4_b200 import _lib as L
    lib = L.load()
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K).cuda()
    A[:, :4] *= 30.                                   # a few outlier columns, like a residual stream
    W = (torch.randn(N, K) / math.sqrt(K)).cuda()
    hi, lo, ldw, inv_q = pack_f16(W)
    bias, rs = torch.randn(N).cuda(), torch.rand(M).cuda() + 0.5
    nout = N // 2 if act else N
    res = torch.randn(M, nout).cuda() if not act else None
    ldc = (nout + 3) // 4 * 4                          # output rows are TMA-stored: 16-byte aligned
    Cc = torch.full((M, ldc), float('nan')).cuda()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    L.check(lib.d4_linear(D4_PREC_F16X3, M, N, K, L.ptr(A), K, L.ptr(hi), ldw, L.ptr(lo), L.ptr(bias), L.ptr(rs * inv_q), L.ptr(res),
                          nout, act, L.ptr(Cc), ldc, stream))
    Cc = Cc[:, :nout]
    torch.cuda.synchronize()
    ref = (A.double() @ W.double().T) * rs.double()[:, None] + bias.double()
    if act:
        x, g = ref[:, 0::2], ref[:, 1::2]
        ref = x * (torch.nn.functional.silu(g) if act == 1 else torch.nn.functional.gelu(g))
    else:
        ref = ref + res.double()
    assert not torch.isnan(Cc).any()
    tol = 1e-5 * max(1.0, K / 256)
    err = (Cc.double() - ref).abs().max().item()
    assert err < tol * max(1.0, ref.abs().max().item()), f'max abs err {err}'


# ------------------------------------------------------------------------------------------------ the engine in f16x3 mode
# DynamicsWorldModel(precision='f16x3'): every transformer GEMM the fp16 kernel takes (M > 128 rows, K a multiple of 32) runs on
# it, the rest - heads, d4_learn, odd shapes - on 3xTF32.  Same oracle, same tolerances as the tf32x3 mode in test_gpu_parity.py.

@pytest.mark.parametrize('name', ['config2_mnist', 'config4_256px'])
def test_f16x3_engine_matches_oracle(name):
    import test_gpu_parity as G
    from dreamer4_b200 import DynamicsWorldModel
    from oracle import dreamer4_oracle as O
    kwargs = G.BASELINE_MODELS[name]
    torch.manual_seed(21)
    model = DynamicsWorldModel(**kwargs, precision='f16x3')
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith('gamma') or n.endswith('norm.weight') or n.endswith('norm_context.weight'):
                p.add_(torch.randn_like(p) * 0.1)
            if 'unembed' in n or n.endswith('queries') or 'learned_embed' in n or n == 'register_tokens':
                p.mul_(30.)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.cuda()
    ocfg = O.config_from_reference_kwargs(**kwargs)
    T, B = 3, 20                                      # B * S = 300 rows: the pair kernels are the ones dispatched
    noise = G.make_noise(model.cfg, T, B, seed=17)
    ref = O.generate(sd, ocfg, T, B, noise=O.InjectedNoise(noise['latent'], noise['action_uniform'], noise['terminal_uniform']))
    exp, tc = model.generate(T, batch_size=B, return_rewards_per_frame=True, return_agent_actions=True,
                             return_log_probs_and_values=True, return_time_cache=True, noise=G.to_cuda(noise))
    ref_kv = torch.stack([torch.stack(layer) for layer in ref.kv_cache])
    G.compare_experience(exp, ref, tc.main.next_kv_cache, ref_kv, TOL=dict(atol=2e-4, rtol=2e-4), LOGIT_TOL=dict(atol=4e-4, rtol=2e-4))
    keys = [k for k in sd if k.startswith(('policy_head.', 'value_head.')) or k == 'action_embedder.discrete_action_unembed']
    sdg = {k: (v.clone().requires_grad_(True) if k in keys else v) for k, v in sd.items()}
    rpl, rvl, _ = O.learn_from_experience(sdg, ocfg, ref)
    pl, vl = model.learn_from_experience(exp)
    torch.testing.assert_close(pl.detach().cpu(), rpl.detach(), atol=2e-6, rtol=1e-4)
    torch.testing.assert_close(vl.detach().cpu(), rvl.detach(), atol=2e-6, rtol=1e-4)
