"""Mirror of the reference's models/analysis_17.py (Analysis_net_17): same sub-module names, shapes
and init order (analysis_17.py:12-25); forward (analysis_17.py:32-39) runs on libb2c kernels."""
import math

import torch
import torch.nn as nn

from .. import functional as B2F
from .GDN import GDN


class Analysis_net_17(nn.Module):
    """conv 9x9/4 -> GDN -> conv 5x5/2 -> GDN -> conv 5x5/2 (no bias)."""

    def __init__(self, out_channel_N=192):
        super().__init__()
        N = out_channel_N
        # real torch modules are the parameter containers: their default init consumes the global RNG
        # exactly as the reference's constructor does, so seeded weights are identical.
        self.conv1 = nn.Conv2d(3, N, 9, stride=4, padding=4)
        torch.nn.init.xavier_normal_(self.conv1.weight.data, math.sqrt(2 * (3 + N) / 6))
        torch.nn.init.constant_(self.conv1.bias.data, 0.01)
        self.gdn1 = GDN(N)
        self.conv2 = nn.Conv2d(N, N, 5, stride=2, padding=2)
        torch.nn.init.xavier_normal_(self.conv2.weight.data, math.sqrt(2))
        torch.nn.init.constant_(self.conv2.bias.data, 0.01)
        self.gdn2 = GDN(N)
        self.conv3 = nn.Conv2d(N, N, 5, stride=2, padding=2, bias=False)
        torch.nn.init.xavier_normal_(self.conv3.weight.data, math.sqrt(2))

    def forward(self, x):
        x = self.gdn1(B2F.conv2d(x, self.conv1.weight, self.conv1.bias, 4, 4))
        x = self.gdn2(B2F.conv2d(x, self.conv2.weight, self.conv2.bias, 2, 2))
        return B2F.conv2d(x, self.conv3.weight, None, 2, 2)
