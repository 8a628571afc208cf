"""Mini `einx` — named-axis elementwise ops, hot-path subset only (test infrastructure).

Supports the patterns /root/reference/dreamer4/dreamer4.py uses on the default path:
'a b, b -> a b', '... na, na', '... h n d, h d', '1 d, b t d', 'b ... d, b d'.
Semantics follow einx >= 0.3: every input is broadcast to the output expression by
axis name; with no '->' the output is the input expression that contains all axes.
"""
import torch

def _parse(expr):
    return expr.strip().split()

def _elementwise(op, pattern, *tensors):
    if '->' in pattern:
        lhs, out = pattern.split('->')
        out_axes = _parse(out)
    else:
        lhs, out_axes = pattern, None
    in_axes = [_parse(e) for e in lhs.split(',')]
    assert len(in_axes) == len(tensors), (pattern, len(tensors))

    def named(ax):
        return [a for a in ax if a != '...' and not a.isdigit()]

    if out_axes is None:
        all_named = set(a for ax in in_axes for a in named(ax))
        cands = [ax for ax in in_axes if set(named(ax)) >= all_named]
        # prefer the candidate carrying the ellipsis / most axes
        cands.sort(key = lambda ax: (('...' in ax), len(ax)), reverse = True)
        assert cands, f'cannot infer output for {pattern}'
        out_axes = cands[0]

    # resolve ellipsis rank from the tensors
    ell_rank = 0
    for ax, t in zip(in_axes, tensors):
        if '...' in ax:
            ell_rank = max(ell_rank, t.ndim - (len(ax) - 1))

    def expand_axes(ax):
        res = []
        for a in ax:
            if a == '...':
                res.extend(f'_e{i}' for i in range(ell_rank))
            else:
                res.append(a)
        return res

    out_full = expand_axes(out_axes)
    views = []
    for ax, t in zip(in_axes, tensors):
        full = expand_axes(ax)
        assert len(full) == t.ndim, (pattern, full, tuple(t.shape))
        # drop literal-1 axes, then place the remaining by name
        keep = [i for i, a in enumerate(full) if not a.isdigit()]
        for i, a in enumerate(full):
            if a.isdigit():
                assert t.shape[i] == int(a)
        t = t.reshape([t.shape[i] for i in keep])
        names = [full[i] for i in keep]
        order = [names.index(a) for a in out_full if a in names]
        t = t.permute(order)
        shape, j = [], 0
        for a in out_full:
            if a in names:
                shape.append(t.shape[j]); j += 1
            else:
                shape.append(1)
        views.append(t.reshape(shape))
    res = views[0]
    for v in views[1:]:
        res = op(res, v)
    # literal-digit output axes are already size-1 dims
    return res

def add(pattern, *ts):            return _elementwise(torch.add, pattern, *ts)
def multiply(pattern, *ts):       return _elementwise(torch.mul, pattern, *ts)
def equal(pattern, *ts):          return _elementwise(torch.eq, pattern, *ts)
def logical_and(pattern, *ts):    return _elementwise(torch.logical_and, pattern, *ts)
def greater_equal(pattern, *ts):  return _elementwise(torch.ge, pattern, *ts)
def less(pattern, *ts):           return _elementwise(torch.lt, pattern, *ts)
def subtract(pattern, *ts):       return _elementwise(torch.sub, pattern, *ts)

def where(pattern, cond, a, b):
    raise NotImplementedError('einx.where is off the hot path')

def dot(pattern, *ts):
    raise NotImplementedError('einx.dot is off the hot path')
