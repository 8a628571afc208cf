"""Model-level accuracy of the operand splits at the config-4 architecture, on the CPU: one transformer pass of the engine dataflow
(tests/engine_emulator.py) with every tensor-core GEMM replaced by its exact operand-split emulation, against the same pass in fp64.
    python scripts/f16x3_model_emulation.py > profiles/r1_f16x3_model_emulation.txt
modes: fp32 (plain), tf32x3, f16x3 (rows pre-scaled where the engine scales them), f16x3_nopow2 (never), f16x3_allpow2 (every GEMM)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from dreamer4_b200 import DynamicsWorldModel
from dreamer4_b200.packing import pack
from engine_emulator import emulate_pass, split_packed
import bench
kw = bench.WORKLOADS['config4']['model']
torch.manual_seed(0)
model = DynamicsWorldModel(**kw); cfg = model.cfg
P = pack(model.state_dict(), cfg, torch.device('cpu'))
P64 = {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k,v in P.items()}
B=2
x = torch.randn(B, cfg.num_latent_tokens, cfg.dim_latent)
res={}
for mode in ('fp32','tf32x3','f16x3','f16x3_nopow2','f16x3_allpow2'):
    Pm = P if mode=='fp32' else split_packed(P,'f16x3',False) if mode=='f16x3_nopow2' else split_packed(P,'f16x3',True) if mode=='f16x3_allpow2' else split_packed(P, mode)
    res[mode] = emulate_pass(Pm, cfg, x, 48, 4, None, None, 0)[:2]
torch.set_default_dtype(torch.float64)
pr, ar, _ = emulate_pass(P64, cfg, x.double(), 48, 4, None, None, 0)
torch.set_default_dtype(torch.float32)
for mode,(pe,ae) in res.items():
    print(mode, 'pred err %.2e agent err %.2e' % ((pe.double()-pr).abs().max().item(), (ae.double()-ar).abs().max().item()), 'scale', pr.abs().max().item(), ar.abs().max().item())
