"""Mirror of the reference's models/synthesis_17.py (Synthesis_net_17): same sub-module names, shapes
and init order (synthesis_17.py:13-25); forward (synthesis_17.py:27-31) runs on libb2c kernels."""
import math

import torch
import torch.nn as nn

from .. import functional as B2F
from .GDN import GDN


class Synthesis_net_17(nn.Module):
    """deconv 5x5/2 -> IGDN -> deconv 5x5/2 -> IGDN -> deconv 9x9/4 (N -> 3)."""

    def __init__(self, out_channel_N=192):
        super().__init__()
        N = out_channel_N
        self.deconv1 = nn.ConvTranspose2d(N, N, 5, stride=2, padding=2, output_padding=1)
        torch.nn.init.xavier_normal_(self.deconv1.weight.data, math.sqrt(2 * 1))
        torch.nn.init.constant_(self.deconv1.bias.data, 0.01)
        self.igdn1 = GDN(N, inverse=True)
        self.deconv2 = nn.ConvTranspose2d(N, N, 5, stride=2, padding=2, output_padding=1)
        torch.nn.init.xavier_normal_(self.deconv2.weight.data, math.sqrt(2 * 1))
        torch.nn.init.constant_(self.deconv2.bias.data, 0.01)
        self.igdn2 = GDN(N, inverse=True)
        self.deconv3 = nn.ConvTranspose2d(N, 3, 9, stride=4, padding=4, output_padding=3)
        torch.nn.init.xavier_normal_(self.deconv3.weight.data, math.sqrt(2 * 1))
        torch.nn.init.constant_(self.deconv3.bias.data, 0.01)

    def forward(self, x):
        x = self.igdn1(B2F.conv_transpose2d(x, self.deconv1.weight, self.deconv1.bias, 2, 2, 1))
        x = self.igdn2(B2F.conv_transpose2d(x, self.deconv2.weight, self.deconv2.bias, 2, 2, 1))
        return B2F.conv_transpose2d(x, self.deconv3.weight, self.deconv3.bias, 4, 4, 3)
