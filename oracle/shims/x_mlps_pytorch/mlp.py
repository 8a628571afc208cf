"""`x_mlps_pytorch.MLP` restated (test infrastructure; parity unpinned)."""
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
                layer = nn.Sequential(layer, activation)
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
