from torch import nn

Identity = nn.Identity

def Sequential(*modules):
    return nn.Sequential(*[m for m in modules if m is not None])
