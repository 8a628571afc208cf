"""`hl_gauss_pytorch.HLGaussLoss` restated, hot-path subset (test infrastructure; PARITY UNPINNED).
Call sites: /root/reference/dreamer4/dreamer4.py:1059-1105."""
from math import sqrt
import torch
from torch import nn

class HLGaussLoss(nn.Module):
    def __init__(self, min_value, max_value, num_bins, sigma = None, sigma_to_bin_ratio = None,
                 eps = 1e-10, clamp_to_range = False, min_max_value_on_bin_center = False):
        super().__init__()
        self.eps = eps
        self.num_bins = num_bins
        self.clamp_to_range = clamp_to_range
        if min_max_value_on_bin_center:
            adjust = (max_value - min_value) / ((num_bins - 1) * 2)
            min_value, max_value = min_value - adjust, max_value + adjust
        self.min_value, self.max_value = min_value, max_value
        support = torch.linspace(min_value, max_value, num_bins + 1).float()
        bin_size = (max_value - min_value) / num_bins
        if sigma is None:
            sigma = sigma_to_bin_ratio * bin_size
        self.sigma = sigma
        self.sigma_times_sqrt_two = sqrt(2.) * sigma
        self.register_buffer('support', support, persistent = False)
        self.register_buffer('centers', (support[:-1] + support[1:]) / 2, persistent = False)

    def transform_to_probs(self, target, eps = None):
        eps = self.eps if eps is None else eps
        if self.clamp_to_range:
            target = target.clamp(self.min_value, self.max_value)
        cdf = torch.special.erf((self.support - target[..., None]) / self.sigma_times_sqrt_two)
        z = cdf[..., -1] - cdf[..., 0]
        probs = cdf[..., 1:] - cdf[..., :-1]
        return probs / z.clamp(min = eps)[..., None]

    def transform_from_probs(self, probs):
        return (probs * self.centers).sum(dim = -1)

    def transform_from_logits(self, logits):
        return self.transform_from_probs(logits.softmax(dim = -1))
