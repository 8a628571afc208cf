import torch
import torch.nn.functional as F
from torch import nn

class ReluSquared(nn.Module):
    def forward(self, x):
        return F.relu(x) ** 2

class SugarBSiLU(nn.Module):
    def forward(self, x):
        # straight-through surrogate; forward value is relu
        return F.relu(x)
