"""`discrete_continuous_embed_readout` discrete subset restated (test infrastructure; PARITY UNPINNED).
Call sites: /root/reference/dreamer4/dreamer4.py:1375-1376, 1422-1426, 1478-1481.
Sampling is gumbel-argmax in the exact form the reference itself uses for its own helper
(dreamer4.py:473-497): argmax(logits / max(T, 1e-10) - log(-log(u))), u ~ rand_like(logits),
log(t) = t.clamp(min = 1e-20).log(); one `rand_like` per action type, in order."""
import torch
from torch import nn

def _log(t, eps = 1e-20):
    return t.clamp(min = eps).log()

class MultiCategorical:
    def __init__(self, logits, use_parallel_multi_discrete = None):
        self.logits = list(logits) if isinstance(logits, (tuple, list)) else [logits]

    def sample(self, temperature = 1., eps = 1e-10):
        out = []
        for l in self.logits:
            noise = torch.rand_like(l)
            g = -_log(-_log(noise))
            out.append(((l / max(temperature, eps)) + g).argmax(dim = -1))
        return torch.stack(out, dim = -1)

    def log_prob(self, targets):
        out = []
        for i, l in enumerate(self.logits):
            lp = l.log_softmax(dim = -1)
            tgt = targets[..., i]
            tgt = tgt.expand(lp.shape[:-1])
            # the world-model training branch hands in -1 sentinels at positions it masks out afterwards (dreamer4.py:7525, 7567): any
            # in-range index does there, the value is discarded
            out.append(lp.gather(-1, tgt.clamp(min = 0)[..., None])[..., 0])
        return torch.stack(out, dim = -1)

    def entropy(self):
        out = []
        for l in self.logits:
            lp = l.log_softmax(dim = -1)
            out.append(-(lp.exp() * lp).sum(dim = -1))
        return torch.stack(out, dim = -1)

    def kl_div(self, other, keep_num_actions_dim = False):
        out = []
        for l, o in zip(self.logits, other.logits):
            lp, op = l.log_softmax(dim = -1), o.log_softmax(dim = -1)
            out.append((lp.exp() * (lp - op)).sum(dim = -1))
        out = torch.stack(out, dim = -1)
        return out if keep_num_actions_dim else out.sum(dim = -1)

class Readout(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        raise NotImplementedError('continuous actions are a "next" row (SURVEY.md 8f)')

class BetaDist(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()

def rescale(t, *args, **kwargs):
    raise NotImplementedError
