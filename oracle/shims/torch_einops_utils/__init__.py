"""Hot-path subset of `torch-einops-utils` (test infrastructure; semantics restated, parity unpinned)."""
import torch
import torch.nn.functional as F
from torch.utils._pytree import tree_flatten, tree_unflatten, tree_map

def exists(v): return v is not None

def maybe(fn):
    def inner(t, *args, **kwargs):
        if not exists(t) or not exists(fn):
            return t
        return fn(t, *args, **kwargs)
    return inner

def pad_right_ndim_to(t, ndim):
    return t.reshape(*t.shape, *((1,) * max(0, ndim - t.ndim)))

def align_dims_left(tensors):
    ndim = max(t.ndim for t in tensors)
    return tuple(pad_right_ndim_to(t, ndim) for t in tensors)

def pad_at_dim(t, pad, dim = -1, value = 0.):
    dims_from_right = (- dim - 1) if dim < 0 else (t.ndim - dim - 1)
    zeros = (0, 0) * dims_from_right
    return F.pad(t, (*zeros, *pad), value = value)

def pad_left_at_dim(t, pad, dim = -1, value = 0.):
    return pad_at_dim(t, (pad, 0), dim = dim, value = value)

def pad_right_at_dim(t, pad, dim = -1, value = 0.):
    return pad_at_dim(t, (0, pad), dim = dim, value = value)

def pad_right_at_dim_to(t, length, dim = -1, value = 0.):
    curr = t.shape[dim]
    if curr >= length:
        return t
    return pad_right_at_dim(t, length - curr, dim = dim, value = value)

def lens_to_mask(lens, max_len = None):
    if not exists(max_len):
        max_len = int(lens.amax().item())
    seq = torch.arange(max_len, device = lens.device)
    return seq < lens[..., None]

def shift_right(t, amount = 1, dim = 1, value = 0.):
    t = pad_at_dim(t, (amount, -amount), dim = dim, value = value)
    return t

def masked_mean(t, mask = None, dim = None, eps = 1e-5):
    if not exists(mask):
        return t.mean(dim = dim) if exists(dim) else t.mean()
    if mask.ndim < t.ndim:
        mask = pad_right_ndim_to(mask, t.ndim)
    mask = mask.expand_as(t)
    if not exists(dim):
        return t[mask].mean() if mask.any() else t[mask].sum()
    num = (t * mask).sum(dim = dim)
    den = mask.sum(dim = dim)
    return num / den.clamp(min = eps)

def repeat_interleave_to_match(t, target):
    return t.repeat_interleave(target.shape[0] // t.shape[0], dim = 0)

def safe_stack(tensors, dim = 0):
    tensors = [t for t in tensors if exists(t)]
    if len(tensors) == 0:
        return None
    return torch.stack(tensors, dim = dim)

def safe_cat(tensors, dim = 0):
    tensors = [t for t in tensors if exists(t)]
    if len(tensors) == 0:
        return None
    if len(tensors) == 1:
        return tensors[0]
    return torch.cat(tensors, dim = dim)

def tree_flatten_with_inverse(tree):
    flat, spec = tree_flatten(tree)
    def inverse(out):
        return tree_unflatten(list(out), spec)
    return flat, inverse

def tree_map_tensor(fn, tree):
    return tree_map(lambda t: fn(t) if torch.is_tensor(t) else t, tree)
