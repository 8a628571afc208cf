"""Golden vectors for the stand-alone AxialSpaceTimeTransformer (reference dreamer4.py:2762-3267; exported at dreamer4/__init__.py):
a 3-frame forward of the reference's own module with its intermediates (time-KV cache).  Build-container only; fixture committed.

    python oracle/make_golden_axial.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_import import import_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden', 'cache', 'axial_transformer.pt')


def main():
    ref = import_reference()
    torch.manual_seed(5)
    kw = dict(dim=32, depth=3, attn_heads=2, attn_dim_head=16, time_block_every=2, num_special_tokens=2, final_norm=True)
    m = ref.AxialSpaceTimeTransformer(**kw).eval()
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith('gamma') or n.endswith('norm.weight') or n.endswith('norm_context.weight'):
                p.add_(torch.randn_like(p) * 0.1)
        tokens = torch.randn(2, 3, 7, 32)
        out, inter = m(tokens, return_intermediates=True)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    torch.save(dict(kwargs=kw, state_dict=sd, tokens=tokens, out=out, kv=inter.next_kv_cache, token_count=inter.token_count,
                    torch_version=torch.__version__), OUT)
    print('wrote', OUT)


if __name__ == '__main__':
    main()
