"""CPU checks of the host-side logic: state_dict key parity with the reference, the packing algebra of
dreamer4_b200/packing.py (through tests/engine_emulator.py, which mirrors engine.cu's dataflow) against the oracle,
and that the C-ABI library loads and exports every symbol include/d4b200.h declares."""
import glob
import math
import os
import re

import pytest
import torch

from dreamer4_b200 import DynamicsWorldModel
from dreamer4_b200.packing import f16_split, pack, tf32_split
from oracle import dreamer4_oracle as O
from engine_emulator import emulate_pass, split_packed

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden', '*.pt')))
IDS = [os.path.basename(p)[:-3] for p in GOLDEN]


def load(path):
    return torch.load(path, map_location='cpu', weights_only=False)


@pytest.mark.parametrize('path', GOLDEN, ids=IDS)
def test_state_dict_keys_match_reference(path):
    fx = load(path)
    model = DynamicsWorldModel(**fx['model_kwargs'])
    model.load_state_dict(fx['state_dict'], strict=True)
    for k, v in fx['state_dict'].items():
        assert model.state_dict()[k].shape == v.shape, k


@pytest.mark.parametrize('path', GOLDEN, ids=IDS)
def test_packed_dataflow_matches_oracle(path):
    """Three frames (clean passes commit the KV cache) of the emulated engine dataflow vs oracle.forward_step."""
    fx = load(path)
    sd = fx['state_dict']
    model = DynamicsWorldModel(**fx['model_kwargs'])
    cfg = model.cfg
    ocfg = O.config_from_reference_kwargs(**fx['model_kwargs'])
    P = pack(sd, cfg, torch.device('cpu'))
    torch.manual_seed(0)
    B = 3
    kv_o, kv_e, prev = None, None, None
    for t in range(3):
        x = torch.randn(B, cfg.num_latent_tokens, cfg.dim_latent)
        for signal in (0, 48, 63):
            po, ao, nko = O.forward_step(sd, ocfg, x, signal, 4, prev, kv_o, t)
            pe, ae, nke = emulate_pass(P, cfg, x, signal, 4, prev, kv_e, t)
            torch.testing.assert_close(pe, po, atol=2e-5, rtol=1e-4)
            torch.testing.assert_close(ae, ao, atol=2e-5, rtol=1e-4)
        kv_o, kv_e = nko, nke
        for (ko, vo), (ke, ve) in zip(kv_o, kv_e):
            torch.testing.assert_close(ke, ko, atol=2e-5, rtol=1e-4)
            torch.testing.assert_close(ve, vo, atol=2e-5, rtol=1e-4)
        prev = torch.stack([torch.randint(0, n, (B,)) for n in cfg.num_discrete_actions], dim=-1)


@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p)[:-3] for p in GOLDEN])
def test_parameter_groups_match_reference(path):
    """muon_parameters / policy_head_parameters / value_head_parameters name the same tensors as the reference's (D4:5335-5363);
    the policy group is the policy MLP + the action UNembeddings (a set in the reference, 1249-1250), not the action embedding."""
    fx = torch.load(path, map_location='cpu', weights_only=False)
    model = DynamicsWorldModel(**fx['model_kwargs'])
    names = {id(p): n for n, p in model.named_parameters()}
    for fn, want in fx['out']['param_groups'].items():
        got = [names[id(p)] for p in getattr(model, fn)()]
        assert sorted(got) == sorted(want), fn
        if fn != 'policy_head_parameters':
            assert got == want, fn


def test_save_and_init_and_load_round_trip(tmp_path):
    """.save / .load / .init_and_load of the reference's @save_load (D4:4660; tests/test_dreamer.py:2243-2247 of the reference):
    the constructor arguments travel with the state_dict, so a checkpoint rebuilds the same architecture and weights."""
    fx = torch.load(GOLDEN[1], map_location='cpu', weights_only=False)
    model = DynamicsWorldModel(**fx['model_kwargs'])
    path = tmp_path / 'world.pt'
    model.save(path)
    with pytest.raises(AssertionError):
        model.save(path, overwrite=False)
    clone = DynamicsWorldModel.init_and_load(path)
    assert clone.cfg == model.cfg
    sd, sd2 = model.state_dict(), clone.state_dict()
    assert list(sd) == list(sd2) and all(torch.equal(sd[k], sd2[k]) for k in sd)
    other = DynamicsWorldModel(**fx['model_kwargs'])
    other.load(path)
    assert all(torch.equal(sd[k], v) for k, v in other.state_dict().items())


def test_f16_split_words():
    """f16x3 packing (experimental mode): q is a power of two bringing rms(q w) into [2^-0.5, 2^0.5], hi + lo reproduces q w to
    2^-22 relative (or fp16's subnormal spacing), both words finite for weights of any sane scale."""
    torch.manual_seed(1)
    for scale in (1e-4, 0.02, 1.0, 300.0):
        w = torch.randn(96, 512) * scale
        hi, lo, inv_q = f16_split(w)
        assert hi.dtype == lo.dtype == torch.float16 and torch.isfinite(hi).all() and torch.isfinite(lo).all()
        q = 1.0 / inv_q
        assert math.log2(q) == round(math.log2(q))
        rms = (w * q).pow(2).mean().sqrt().item()
        assert 2 ** -0.51 <= rms <= 2 ** 0.51
        err = (hi.double() + lo.double() - w.double() * q).abs()
        # |lo| <= 2^-11 |x| is itself rounded to 11 bits (2^-22 |x|) or, where it falls below fp16's normal range, to the
        # subnormal grid (spacing 2^-24)
        assert (err <= torch.maximum((w.double() * q).abs() * 2.0 ** -22, torch.tensor(2.0 ** -25, dtype=torch.float64))).all()


@pytest.mark.parametrize('path', GOLDEN, ids=IDS)
def test_f16x3_dataflow_keeps_fp32_accuracy(path):
    """The engine dataflow with every tensor-core GEMM replaced by its operand-split emulation (tests/engine_emulator.py:
    SplitWeight) against the same dataflow in fp64: the fp16 3-term split with power-of-two pre-scales (f16x3, experimental)
    stays within 2x of what plain fp32 arithmetic loses - the bar the 3xTF32 split meets - on pred and agent embedding."""
    fx = load(path)
    model = DynamicsWorldModel(**fx['model_kwargs'])
    cfg = model.cfg
    P = pack(fx['state_dict'], cfg, torch.device('cpu'))
    P64 = {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in P.items()}
    torch.manual_seed(0)
    x = torch.randn(4, cfg.num_latent_tokens, cfg.dim_latent)
    torch.set_default_dtype(torch.float64)
    try:
        pr, ar, _ = emulate_pass(P64, cfg, x.double(), 48, 4, None, None, 0)
    finally:
        torch.set_default_dtype(torch.float32)
    err = {}
    for mode in ('fp32', 'tf32x3', 'f16x3'):
        pe, ae, _ = emulate_pass(P if mode == 'fp32' else split_packed(P, mode), cfg, x, 48, 4, None, None, 0)
        err[mode] = max((pe.double() - pr).abs().max().item() / pr.abs().max().item(), (ae.double() - ar).abs().max().item() / ar.abs().max().item())
    assert err['f16x3'] <= 2 * err['fp32'] + 1e-7, err
    assert err['tf32x3'] <= 2 * err['fp32'] + 1e-7, err


def test_tf32_split_round_to_nearest():
    """w = hi + lo with both words TF32-representable (so the tensor core's operand truncation is a no-op), hi the NEAREST
    TF32 value (|lo| <= 2^-11 |w|: half a TF32 ulp) and a residual of at most 2^-23 |w|: an fp32 ulp."""
    w = torch.randn(1000) * torch.logspace(-6, 6, 1000)
    hi, lo = tf32_split(w)
    assert torch.all((hi.view(torch.int32) & 0x1FFF) == 0) and torch.all((lo.view(torch.int32) & 0x1FFF) == 0)
    assert torch.all(lo.abs() <= w.abs() * 2 ** -11 * (1 + 2 ** -10))
    assert torch.all((w.double() - hi.double() - lo.double()).abs() <= w.abs().double() * 2 ** -23)


def test_library_exports_every_declared_symbol():
    from dreamer4_b200 import _lib
    header = open(os.path.join(os.path.dirname(os.path.dirname(__file__)), 'include', 'd4b200.h')).read()
    declared = set(re.findall(r'\b(d4_[a-z0-9_]+)\s*\(', header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.load()          # raises if the .so is missing or a symbol is not exported
    assert lib.d4_version() >= 100
    assert lib.d4_last_error() is not None


def test_integration_md_names_every_entry_point():
    """INTEGRATION.md says, per exported entry point, which reference interface it replaces (file:line): none may be missing."""
    root = os.path.dirname(os.path.dirname(__file__))
    declared = set(re.findall(r'\b(d4_[a-z0-9_]+)\s*\(', open(os.path.join(root, 'include', 'd4b200.h')).read()))
    text = open(os.path.join(root, 'INTEGRATION.md')).read()
    assert not sorted(sym for sym in declared if sym not in text)


def test_ctypes_structs_mirror_the_header(tmp_path):
    """Every field of every struct the C-ABI passes by pointer sits at the offset gcc gives it from include/d4b200.h."""
    import ctypes
    import subprocess
    from dreamer4_b200 import _lib
    root = os.path.dirname(os.path.dirname(__file__))
    structs = {'d4_config': _lib.d4_config, 'd4_frame_io': _lib.d4_frame_io, 'd4_learn_io': _lib.d4_learn_io, 'd4_tf_config': _lib.d4_tf_config}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "d4b200.h"', 'int main(void) {']
    for name, cls in structs.items():
        lines.append(f'  printf("{name} %zu\\n", sizeof({name}));')
        for field, _ in cls._fields_:
            lines.append(f'  printf("{name}.{field} %zu\\n", offsetof({name}, {field}));')
    lines += ['  return 0;', '}']
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'layout'
    subprocess.run(['gcc', '-I', os.path.join(root, 'include'), str(src), '-o', str(exe)], check=True)
    got = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for name, cls in structs.items():
        assert int(got[name]) == ctypes.sizeof(cls), name
        for field, _ in cls._fields_:
            assert int(got[f'{name}.{field}']) == getattr(cls, field).offset, f'{name}.{field}'


def test_sass_is_native_blackwell():
    """The built library is sm_100a code whose hot kernels use the Blackwell units the design claims: tcgen05 MMAs (UTCHMMA, incl.
    the CTA-pair form), TMEM loads (LDTM), TMA tile loads / stores (UTMALDG / UTMASTG), bulk copies (UBLKCP: the K1 ring), and no
    kernel spills registers to local memory.  (Mnemonics per /opt/skills/guides/B200_PROFILING.md.)"""
    import shutil
    import subprocess
    from dreamer4_b200 import _lib
    if not shutil.which('cuobjdump'):
        pytest.skip('cuobjdump not on PATH')
    sass = subprocess.run(['cuobjdump', '-sass', _lib.LIB_PATH], check=True, capture_output=True, text=True).stdout
    assert 'arch = sm_100a' in sass and 'arch = sm_90' not in sass
    for opcode in ('UTCHMMA.2CTA', 'UTCHMMA', 'LDTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'UTCBAR'):
        assert sass.count(opcode) > 0, opcode
    usage = subprocess.run(['cuobjdump', '-res-usage', _lib.LIB_PATH], check=True, capture_output=True, text=True).stdout
    names = re.findall(r'Function (\S+):\n\s*REG:(\d+) STACK:(\d+) SHARED:\d+ LOCAL:(\d+)', usage)
    assert len(names) > 20
    hot = [(n, int(reg), int(local)) for n, reg, stack, local in names if re.search(r'gemm_tc3_kernel|gemm_tc2_kernel|time_attn_bulk_kernel|gemm_f16x3_kernel', n)]
    assert hot and all(local == 0 for _, _, local in hot), [h for h in hot if h[2]]


def test_no_cpu_fallback():
    from dreamer4_b200._lib import D4Error
    model = DynamicsWorldModel(dim=32, dim_latent=8, num_latent_tokens=6, attn_heads=2, attn_dim_head=16, num_discrete_actions=4)
    with pytest.raises(D4Error):
        model.generate(2, batch_size=1)
