"""Plugin registries of the reference (`register_activation`, dreamer4/dreamer4.py:554-576; `register_reward_encoder`, :1107-1117).

The reference looks activation / reward-encoder classes up by name when it builds its torch modules.  On this path the arithmetic
behind those names lives in CUDA kernels, so a registration is recorded (the reference's own scripts and tests may call the hooks)
and the name -> kernel table below says which registered names the native path can execute: the gated feed-forward epilogues
cover `silu` and `gelu`, the head MLPs `silu`, the reward / value codec `hl_gauss`.  Asking a model for any other registered name
raises NotImplementedError at construction - never a silent substitution."""
from torch import nn

ACTIVATIONS = dict(silu=nn.SiLU, relu=nn.ReLU, gelu=nn.GELU)          # relu_squared / sugar_bsilu come from x-mlps in the reference
NATIVE_FF_ACTIVATIONS = ('silu', 'gelu')
NATIVE_MLP_ACTIVATIONS = ('silu',)
REWARD_ENCODERS = dict(hl_gauss='native: hl_gauss_decode_kernel / value_row_kernel', symexp_two_hot=None)
NATIVE_REWARD_ENCODERS = ('hl_gauss',)


def register_activation(name, klass):
    ACTIVATIONS[name] = klass


def get_activation(act):
    if isinstance(act, str):
        assert act in ACTIVATIONS, f'activation {act} not found in {list(ACTIVATIONS.keys())}'
        return ACTIVATIONS[act]()
    return act


def register_reward_encoder(name, klass):
    REWARD_ENCODERS[name] = klass


def get_reward_encoder_klass(name):
    assert name in REWARD_ENCODERS, f'unknown reward encoder type {name}'
    return REWARD_ENCODERS[name]
