"""Mirror of the reference's models/analysis_17_new.py (Analysis_net_17_new): the Balle-2017 analysis transform
followed by sigmoid and the straight-through binarizer; returns (code, sigmoid) (analysis_17_new.py:32-38).
Same sub-module names and init order as the reference (analysis_17_new.py:12-24)."""
import math

import torch
import torch.nn as nn

from .. import functional as B2F
from .GDN import GDN
from .binarizer import Lambda, bin_values


class Analysis_net_17_new(nn.Module):
    def __init__(self, out_channel_N=192):
        super().__init__()
        N = out_channel_N
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
        self.sigmoid = nn.Sigmoid()               # parameter-free members kept for attribute parity
        self.binarize = Lambda(bin_values)

    def pre_sigmoid(self, x):
        x = self.gdn1(B2F.conv2d(x, self.conv1.weight, self.conv1.bias, 4, 4))
        x = self.gdn2(B2F.conv2d(x, self.conv2.weight, self.conv2.bias, 2, 2))
        return B2F.conv2d(x, self.conv3.weight, None, 2, 2)

    def forward(self, x):
        return B2F.sigmoid_binarize(self.pre_sigmoid(x))       # (code, x) - analysis_17_new.py:36-38
