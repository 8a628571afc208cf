"""`assoc_scan.AssocScan` restated (test infrastructure; PARITY UNPINNED).
out_t = inputs_t + gates_t * out_{t-1}; reverse=True runs from the right (dreamer4.py:1594-1596)."""
import torch
from torch import nn

class AssocScan(nn.Module):
    def __init__(self, reverse = False, use_accelerated = False):
        super().__init__()
        self.reverse = reverse

    def forward(self, gates, inputs, prev = None):
        if self.reverse:
            gates, inputs = gates.flip(-1), inputs.flip(-1)
        out = torch.empty_like(inputs)
        acc = torch.zeros_like(inputs[..., 0]) if prev is None else prev
        for t in range(inputs.shape[-1]):
            acc = inputs[..., t] + gates[..., t] * acc
            out[..., t] = acc
        return out.flip(-1) if self.reverse else out
