"""Mirror of the reference's models/binarizer.py (Binarizer, bin_values, Lambda): an element-wise threshold at 0.5
with a straight-through gradient (binarizer.py:7-27).  The fused sigmoid + binarizer of Analysis_net_17_new runs in
one libb2c kernel (functional.sigmoid_binarize); these stand-alone forms exist so `from models.binarizer import *`
keeps resolving and compose with any CUDA tensor."""
import torch
from torch import nn


class Binarizer(torch.autograd.Function):
    """binarizer.py:7-27: forward (i > 0.5).float(), backward identity"""

    @staticmethod
    def forward(ctx, i):
        return (i > 0.5).float()

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output


def bin_values(x):
    return Binarizer.apply(x)


class Lambda(nn.Module):
    def __init__(self, func):
        super().__init__()
        self.func = func

    def forward(self, x):
        return self.func(x)
