"""Mirror of the reference's models/bitEstimator.py: Bitparm / BitEstimator hold the same parameters
(h, b, a of shape [1,C,1,1], N(0,0.01) init, bitEstimator.py:13-18).  ``Bitparm.forward`` (bitEstimator.py:20-25) and
``BitEstimator.forward(x)`` (the factorized CDF, bitEstimator.py:38-42) run libb2c's ``b2c_bitparm_fwd/bwd`` with
autograd, for callers that evaluate the CDF themselves (model.py:71-72 pattern, entropy-coder tables);
ImageCompressor uses the fused quantize+bits kernel instead of calling the chain twice."""
import torch
import torch.nn as nn

from .. import functional as B2F


class Bitparm(nn.Module):
    def __init__(self, channel, final=False):
        super().__init__()
        self.final = final
        self.h = nn.Parameter(torch.nn.init.normal_(torch.empty(channel).view(1, -1, 1, 1), 0, 0.01))
        self.b = nn.Parameter(torch.nn.init.normal_(torch.empty(channel).view(1, -1, 1, 1), 0, 0.01))
        if not final:
            self.a = nn.Parameter(torch.nn.init.normal_(torch.empty(channel).view(1, -1, 1, 1), 0, 0.01))
        else:
            self.a = None

    def forward(self, x):
        """bitEstimator.py:20-25"""
        return B2F.bitparm(x, self.h, self.b, None if self.final else self.a)


class BitEstimator(nn.Module):
    def __init__(self, channel):
        super().__init__()
        self.f1 = Bitparm(channel)
        self.f2 = Bitparm(channel)
        self.f3 = Bitparm(channel)
        self.f4 = Bitparm(channel, True)

    def forward(self, x):
        """bitEstimator.py:38-42: CDF = f4(f3(f2(f1(x))))"""
        return self.f4(self.f3(self.f2(self.f1(x))))

    def param_list(self):
        """the 11 per-channel vectors in libb2c order (include/b2c.h)."""
        return [self.f1.h, self.f1.b, self.f1.a, self.f2.h, self.f2.b, self.f2.a,
                self.f3.h, self.f3.b, self.f3.a, self.f4.h, self.f4.b]

    def bits(self, y, noise=None):
        """fused quantizer + bit count: (z, total_bits, image_bits, plane_bits)."""
        return B2F.quant_entropy(y, noise, self.param_list())
