"""`x_mlps_pytorch.normed_mlp` restated (test infrastructure; PARITY UNPINNED).

Believed structure (x-mlps-pytorch >= 0.3.1): every non-final layer is
Linear -> LayerNorm(dim_out) -> activation, the final layer is a bare Linear;
`create_mlp(dim, depth, dim_in, dim_out)` builds dims (dim_in, dim x (depth + 1), dim_out).
The reference call sites are /root/reference/dreamer4/dreamer4.py:4950-4956, 5083-5089, 5095-5101.
dreamer4_b200 mirrors exactly this structure (state-dict keys `layers.{i}.0.*` Linear,
`layers.{i}.1.*` LayerNorm) so the oracle, the shim and the product agree by construction.
"""
from torch import nn

class MLP(nn.Module):
    def __init__(self, *dims, activation = nn.ReLU(), bias = True, activate_last = False):
        super().__init__()
        assert len(dims) > 1
        pairs = tuple(zip(dims[:-1], dims[1:]))
        layers = []
        for i, (dim_in, dim_out) in enumerate(pairs, start = 1):
            is_last = i == len(pairs)
            layer = nn.Linear(dim_in, dim_out, bias = bias)
            if not is_last or activate_last:
                layer = nn.Sequential(layer, nn.LayerNorm(dim_out), activation)
            else:
                layer = nn.Sequential(layer)
            layers.append(layer)
        self.layers = nn.ModuleList(layers)

    def forward(self, x):
        for layer in self.layers:
            x = layer(x)
        return x

def create_mlp(dim, depth, *, dim_in = None, dim_out = None, **kwargs):
    dims = (dim,) * (depth + 1)
    if dim_in is not None:
        dims = (dim_in, *dims)
    if dim_out is not None:
        dims = (*dims, dim_out)
    return MLP(*dims, **kwargs)
