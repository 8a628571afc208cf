"""`x_mlps_pytorch.ensemble.Ensemble` restated (test infrastructure; PARITY UNPINNED).
n independently initialised copies of `net`; `forward` stacks all members on dim 0,
`forward_one(x, id)` runs one member (dreamer4.py:5072-5075, 6598)."""
from copy import deepcopy
import torch
from torch import nn

class Ensemble(nn.Module):
    def __init__(self, net, ensemble_size):
        super().__init__()
        nets = []
        for _ in range(ensemble_size):
            member = deepcopy(net)
            for m in member.modules():
                if isinstance(m, nn.Linear):
                    m.reset_parameters()
            nets.append(member)
        self.nets = nn.ModuleList(nets)
        self.ensemble_size = ensemble_size

    def forward_one(self, *args, id = 0, **kwargs):
        return self.nets[id](*args, **kwargs)

    def forward(self, *args, ids = None, **kwargs):
        nets = self.nets if ids is None else [self.nets[i] for i in ids]
        return torch.stack([net(*args, **kwargs) for net in nets])
