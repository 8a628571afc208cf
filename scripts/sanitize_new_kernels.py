"""Small run of the kernels added at the end of round 2 (frame_attn_mma, lp_fused_persist, ln_act_rows_reg, gemm_rowdot, the early residual
loads of gemm_f16x3) for compute-sanitizer:  compute-sanitizer --tool memcheck|racecheck python scripts/sanitize_new_kernels.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from dreamer4_b200 import AxialSpaceTimeTransformer, DynamicsWorldModel  # noqa: E402

torch.manual_seed(0)
# tokenizer-style transformer: 72 tokens per frame (ragged 8-key / 16-query tiles), 5 special tokens, tensor-core attention inside a frame
tf = AxialSpaceTimeTransformer(dim=128, depth=2, attn_heads=2, attn_dim_head=64, time_block_every=2, num_special_tokens=5, precision='tf32x3').cuda()
out = tf(torch.randn(2, 2, 72, 128, device='cuda'))
assert torch.isfinite(out).all()
# world model: 160 dreams (> 148 SMs: persistent space -> latent pool), 256+ head rows would need B >= 256: use 272 for the row-dot kernel
model = DynamicsWorldModel(dim=128, dim_latent=16, num_latent_tokens=16, depth=2, time_block_every=2, attn_heads=2, attn_dim_head=64,
                           num_discrete_actions=4, predict_terminals=False, precision='f16x3').cuda()
exp = model.generate(2, batch_size=272, return_rewards_per_frame=True, return_agent_actions=True, return_log_probs_and_values=True)
pl, vl = model.learn_from_experience(exp)
torch.cuda.synchronize()
print('ok', float(pl), float(vl), tuple(exp.latents.shape))
