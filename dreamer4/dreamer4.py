"""`dreamer4.dreamer4` names used by the reference's scripts and tests (e.g. train_cartpole_with_dynamics_rl.py:49-59:
Experience, Actions, combine_experiences, exists, default, divisible_by, cast_to_tensor), bound to the B200-native classes."""
import torch

from dreamer4_b200.dynamics import DynamicsWorldModel, ModelConfig, default, exists
from dreamer4_b200.experience import (Actions, DynamicsIntermediates, Embeds, Experience, Predictions, TransformerIntermediates,
                                      combine_experiences)
from dreamer4_b200.registry import (ACTIVATIONS, REWARD_ENCODERS, get_activation, get_reward_encoder_klass, register_activation,
                                    register_reward_encoder)
from dreamer4_b200.tokenizer import AxialSpaceTimeTransformer, VideoTokenizer


def divisible_by(num, den):
    return (num % den) == 0


def cast_to_tensor(t, device, dtype=None):          # reference dreamer4.py:351-354
    t = t if torch.is_tensor(t) else torch.tensor(t)
    return (t.to(dtype) if exists(dtype) else t).to(device)


def calc_gae(rewards, values, masks=None, learn_masks=None, gamma=0.99, lam=0.95, use_accelerated=None):
    """lambda-returns by the native reverse scan (gae_kernel; reference dreamer4.py:1566-1600).  CUDA tensors (b, t)."""
    import ctypes as C
    from dreamer4_b200 import _lib
    assert rewards.is_cuda, 'dreamer4 (B200-native) runs on CUDA only'
    B, T = rewards.shape
    ones = torch.ones(B, T, dtype=torch.uint8, device=rewards.device)
    m = ones if masks is None else masks.to(torch.uint8).contiguous()
    lm = ones if learn_masks is None else learn_masks.to(torch.uint8).contiguous()
    r, v = rewards.float().contiguous(), values.float().contiguous()
    out = torch.empty_like(r)
    _lib.check(_lib.load().d4_gae(B, T, _lib.ptr(r), _lib.ptr(v), _lib.ptr(m), _lib.ptr(lm), gamma, lam, _lib.ptr(out),
                                  C.c_void_p(torch.cuda.current_stream(rewards.device).cuda_stream)))
    return out
