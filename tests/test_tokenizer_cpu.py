"""Host side of the video tokenizer (dreamer4_b200/tokenizer.py) on the CPU: the class reproduces the reference's state_dict
layout, and its packed weights + call sequence (emulated in torch by tests/engine_emulator.py, step for step what the host
sends through the C-ABI) reproduce the reference's golden tokenize / decode vectors (tests/golden/tokenizer/*.pt, produced by
the reference's own source) and the oracle.  The CUDA kernels behind those calls are covered by the -m gpu tests."""
import glob
import os

import pytest
import torch

from dreamer4_b200 import VideoTokenizer
from dreamer4_b200.packing import pack_tokenizer
from engine_emulator import emulate_decode, emulate_tokenize

FIX = sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden', 'tokenizer', 'tokenizer_*.pt')))
IDS = [os.path.basename(p)[:-3] for p in FIX]


def load(path):
    return torch.load(path, map_location='cpu', weights_only=False)


@pytest.mark.parametrize('path', FIX, ids=IDS)
def test_state_dict_layout_matches_reference(path):
    fx = load(path)
    tok = VideoTokenizer(**fx['tokenizer_kwargs'])
    sd = tok.state_dict()
    assert set(sd) == set(fx['state_dict']), sorted(set(sd) ^ set(fx['state_dict']))
    for k, v in fx['state_dict'].items():
        assert sd[k].shape == v.shape, k
    tok.load_state_dict(fx['state_dict'], strict=True)


@pytest.mark.parametrize('path', FIX, ids=IDS)
def test_packed_dataflow_reproduces_reference_golden(path):
    fx = load(path)
    tok = VideoTokenizer(**fx['tokenizer_kwargs'])
    tok.load_state_dict(fx['state_dict'], strict=True)
    PK = pack_tokenizer(tok.state_dict(), tok.cfg, torch.device('cpu'))
    latents = emulate_tokenize(PK, tok.cfg, fx['video'])
    torch.testing.assert_close(latents, fx['latents'], atol=2e-5, rtol=1e-4)
    b, c, T, H, W = fx['video'].shape
    torch.manual_seed(fx['decode_seed'])
    noise = torch.randn(b, c, T, H, W)                                    # the draw at reference dreamer4.py:4204
    recon = emulate_decode(PK, tok.cfg, fx['latents'], noise)
    torch.testing.assert_close(recon, fx['recon'], atol=5e-5, rtol=1e-4)


def test_cpu_tokenizer_refuses_to_run():
    """No CPU fallback: tokenize / decode on a CPU model raise instead of computing."""
    from dreamer4_b200._lib import D4Error
    fx = load(FIX[0])
    tok = VideoTokenizer(**fx['tokenizer_kwargs'])
    with pytest.raises(D4Error):
        tok.tokenize(fx['video'])
    with pytest.raises(D4Error):
        tok.decode(fx['latents'])
