"""ctypes binding of include/d4b200.h.  The structures below mirror the header field for field.

There is no CPU fallback: if libd4b200.so is missing or fails to load, importing the ops raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'csrc', 'libd4b200.so')

D4_MAX_ACTION_TYPES = 8
D4_MAX_MLP_LAYERS = 8
PREC = dict(fp32=0, tf32=1, tf32x3=2, f16x3=3)      # f16x3: experimental, see include/d4b200.h

c_float_p = C.POINTER(C.c_float)


class d4_config(C.Structure):
    _fields_ = [
        ('dim', C.c_int32), ('dim_latent', C.c_int32), ('num_latent_tokens', C.c_int32), ('num_spatial_tokens', C.c_int32),
        ('num_register_tokens', C.c_int32), ('depth', C.c_int32), ('time_block_every', C.c_int32),
        ('heads', C.c_int32), ('query_heads', C.c_int32), ('dim_head', C.c_int32),
        ('pool_heads', C.c_int32), ('pool_dim_head', C.c_int32),
        ('ff_inner', C.c_int32), ('ff_inner_pad', C.c_int32), ('ff_act', C.c_int32),
        ('max_steps', C.c_int32),
        ('num_action_types', C.c_int32), ('action_sizes', C.c_int32 * D4_MAX_ACTION_TYPES),
        ('policy_layers', C.c_int32), ('policy_hidden', C.c_int32),
        ('value_layers', C.c_int32), ('value_hidden', C.c_int32),
        ('terminal_layers', C.c_int32), ('terminal_hidden', C.c_int32), ('predict_terminals', C.c_int32),
        ('reward_bins', C.c_int32), ('value_bins', C.c_int32),
        ('num_tasks', C.c_int32),
        ('softclamp', C.c_float),
        ('max_batch', C.c_int32), ('max_time', C.c_int32),
        ('precision', C.c_int32), ('time_attn_variant', C.c_int32),
    ]


class d4_frame_io(C.Structure):
    _fields_ = [
        ('noise_latent', C.c_void_p), ('action_uniform', C.c_void_p), ('terminal_uniform', C.c_void_p),
        ('prev_actions', C.c_void_p), ('pa_stride', C.c_int64), ('tasks', C.c_void_p),
        ('latents', C.c_void_p), ('latents_bs', C.c_int64),
        ('agent_embed', C.c_void_p), ('agent_bs', C.c_int64),
        ('rewards', C.c_void_p), ('rewards_bs', C.c_int64),
        ('values', C.c_void_p), ('values_bs', C.c_int64),
        ('actions', C.c_void_p), ('actions_bs', C.c_int64),
        ('log_probs', C.c_void_p), ('log_probs_bs', C.c_int64),
        ('logits', C.c_void_p), ('logits_bs', C.c_int64),
        ('lens', C.c_void_p), ('terminals', C.c_void_p),
    ]


class d4_tf_config(C.Structure):
    _fields_ = [
        ('dim', C.c_int32), ('depth', C.c_int32), ('time_block_every', C.c_int32),
        ('heads', C.c_int32), ('query_heads', C.c_int32), ('dim_head', C.c_int32), ('pool_heads', C.c_int32), ('pool_dim_head', C.c_int32),
        ('ff_inner', C.c_int32), ('ff_inner_pad', C.c_int32), ('ff_act', C.c_int32),
        ('tokens_per_frame', C.c_int32), ('num_special', C.c_int32), ('final_norm', C.c_int32),
        ('softclamp', C.c_float),
        ('max_batch', C.c_int32), ('max_time', C.c_int32), ('precision', C.c_int32), ('time_attn_variant', C.c_int32),
    ]


_PTR_ARR = C.c_void_p * D4_MAX_MLP_LAYERS


class d4_learn_io(C.Structure):
    _fields_ = [
        ('B', C.c_int32), ('T', C.c_int32),
        ('agent_embed', C.c_void_p), ('rewards', C.c_void_p), ('old_values', C.c_void_p), ('actions', C.c_void_p),
        ('old_log_probs', C.c_void_p), ('lens', C.c_void_p), ('is_truncated', C.c_void_p), ('terminals', C.c_void_p),
        ('gamma', C.c_float), ('lam', C.c_float), ('eps_clip', C.c_float), ('entropy_weight', C.c_float),
        ('delight_temperature', C.c_float), ('zscore_eps', C.c_float),
        ('use_delight_gating', C.c_int32), ('normalize_advantages', C.c_int32),
        ('value_support', C.c_void_p),
        ('value_sigma_sqrt2', C.c_float), ('hl_eps', C.c_float), ('value_lo', C.c_float), ('value_hi', C.c_float),
        ('losses', C.c_void_p), ('returns', C.c_void_p), ('advantages', C.c_void_p),
        ('grad_policy_w', _PTR_ARR), ('grad_policy_b', _PTR_ARR), ('grad_policy_lnw', _PTR_ARR), ('grad_policy_lnb', _PTR_ARR),
        ('grad_unembed', C.c_void_p), ('grad_unembed_ld', C.c_int64),
        ('grad_value_w', _PTR_ARR), ('grad_value_b', _PTR_ARR), ('grad_value_lnw', _PTR_ARR), ('grad_value_lnb', _PTR_ARR),
        ('objective', C.c_int32), ('pmpo_reverse_kl', C.c_int32),
        ('pmpo_pos_to_neg_weight', C.c_float), ('pmpo_kl_div_loss_weight', C.c_float),
        ('old_action_unembeds', C.c_void_p), ('old_action_unembeds_ld', C.c_int64),
        ('returns_ema', C.c_void_p),
    ]


OBJECTIVES = {'ppo': 0, 'spo': 1, 'pmpo': 2}      # D4_OBJECTIVE_* in include/d4b200.h

# every symbol include/d4b200.h declares: (name, restype, argtypes)
_i, _i64, _f, _p = C.c_int, C.c_int64, C.c_float, C.c_void_p
SYMBOLS = {
    'd4_last_error': (C.c_char_p, []),
    'd4_version': (_i, []),
    'd4_launch_count': (_i64, []),
    'd4_ctx_create': (_i, [C.POINTER(d4_config), C.POINTER(_p)]),
    'd4_ctx_destroy': (None, [_p]),
    'd4_set_weight': (_i, [_p, C.c_char_p, _p, _i64]),
    'd4_set_weight_scale': (_i, [_p, C.c_char_p, _f]),
    'd4_bind': (_i, [_p]),
    'd4_workspace_bytes': (_i64, [_p]),
    'd4_kv_bytes': (_i64, [_p]),
    'd4_set_buffers': (_i, [_p, _p, _i64, _p, _i64]),
    'd4_pass': (_i, [_p, _i, _p, _i, _i, _p, _i64, _p, _i, _i, _p, _p, _p]),
    'd4_frame': (_i, [_p, _i, _i, _i, _f, C.POINTER(d4_frame_io), _p]),
    'd4_observe': (_i, [_p, _i, _i, _i, _f, C.POINTER(d4_frame_io), _p]),
    'd4_profile': (_i, [_p, _i]),
    'd4_profile_read': (_i, [_p, C.POINTER(C.c_double)]),
    'd4_time_attn_decode': (_i, [_i, _i, _i, _i, _i, _i, _p, _i64, _p, _p, _p, _p, _p, _p, _f, _i, _i, _p]),
    'd4_linear': (_i, [_i, _i, _i, _i, _p, _i64, _p, _i64, _p, _p, _p, _p, _i64, _i, _p, _i64, _p]),
    'd4_gae': (_i, [_i, _i, _p, _p, _p, _p, _f, _f, _p, _p]),
    'd4_learn_workspace_bytes': (_i64, [_p, _i, _i]),
    'd4_learn': (_i, [_p, C.POINTER(d4_learn_io), _p, _i64, _p]),
    'd4_tf_create': (_i, [C.POINTER(d4_tf_config), C.POINTER(_p)]),
    'd4_tf_step': (_i, [_p, _i, _p, _i, _p, _p]),
    'd4_linear_rows': (_i, [_i, _i, _i, _i, _p, _i64, _i, _i, _i, _p, _i64, _p, _p, _p, _p, _i64, _p]),
    'd4_patchify': (_i, [_i, _i, _i, _i, _i, _p, _i64, _i64, _p, _p]),
    'd4_unpatchify_flow': (_i, [_i, _i, _i, _i, _i, _p, _p, _i64, _i64, _f, _p]),
    'd4_tok_assemble': (_i, [_i, _i, _i, _i, _p, _p, _p, _p, _i64, _i, _p, _p]),
    'd4_tanh_rows': (_i, [_p, _i64, _p]),
    'd4_pass_ex': (_i, [_p, _i, _p, _p, _p, _p, _i64, _p, _i, _i, _p, _p, _p]),
    'd4_head_forward': (_i, [_p, _i, _p, _i, _p, _p]),
    'd4_graph_replays': (_i64, [_p]),
    'd4_debug_set': (_i, [C.c_char_p, _i]),
    'd4_debug_get': (_i64, [_p, C.c_char_p]),
}

_lib = None


class D4Error(RuntimeError):
    pass


def load():
    """Loads libd4b200.so (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise D4Error(f'{LIB_PATH} is not built: run `python -m dreamer4_b200.build` (there is no CPU fallback)')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise D4Error(load().d4_last_error().decode('utf-8', 'replace'))


def ptr(t):
    """Device/host pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())
