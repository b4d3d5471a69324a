"""Mirror of the reference's model.py: ImageCompressor, save_model, load_model with the same
constructor, attribute names, state_dict keys and forward outputs (model.py:18-81), running on
libb2c's sm_100a kernels.  The star-import surface train.py relies on (np, logging, os, nn, F, ...)
is re-exported exactly as model.py:1-15 does."""
import logging  # noqa: F401  (re-exported: train.py:26 uses `logging` through `from model import *`)
import math  # noqa: F401
import os
import sys  # noqa: F401

import numpy as np  # noqa: F401
import torch
import torch.nn as nn
import torch.nn.functional as F  # noqa: F401
import torch.nn.init as init  # noqa: F401
import torch.optim as optim  # noqa: F401
from torch.nn.parameter import Parameter  # noqa: F401
from torch.utils.data import DataLoader  # noqa: F401

from . import functional as B2F
from .models import *  # noqa: F401,F403
from .models import Analysis_net_17, BitEstimator, Synthesis_net_17


def save_model(model, iter, name):
    """model.py:18-19."""
    torch.save(model.state_dict(), os.path.join(name, "iter_{}.pth.tar".format(iter)))


def load_model(model, f):
    """model.py:22-35: partial load filtered by key; the step parsing is disabled upstream -> returns 0."""
    with open(f, 'rb') as fh:
        pretrained_dict = torch.load(fh, map_location="cpu")
        model_dict = model.state_dict()
        pretrained_dict = {k: v for k, v in pretrained_dict.items() if k in model_dict}
        model_dict.update(pretrained_dict)
        model.load_state_dict(model_dict)
    return 0


class ImageCompressor(nn.Module):
    """model.py:38-81.

    forward(x) -> (clipped_recon, latent, bpp), the HEAD return tuple (model.py:80).
    ``return_mode='mse'`` (or env IMAGECOMPRESSOR_RETURN=mse) selects the commented-out variant
    (model.py:81) that train.py:97 unpacks: (clipped_recon, mse_loss, bpp).
    After every forward, ``self.last`` holds per-image statistics (bits, unclamped / clamped SSE).
    """

    def __init__(self, out_channel_N=128, return_mode=None):
        super().__init__()
        self.Encoder = Analysis_net_17(out_channel_N=out_channel_N)
        self.Decoder = Synthesis_net_17(out_channel_N=out_channel_N)
        self.bitEstimator = BitEstimator(channel=out_channel_N)
        self.out_channel_N = out_channel_N
        mode = return_mode or os.environ.get("IMAGECOMPRESSOR_RETURN", "latent")
        if mode not in ("latent", "mse"):
            raise ValueError("return_mode must be 'latent' (model.py:80) or 'mse' (model.py:81)")
        self.return_mode = mode
        self.noise_override = None      # test hook: inject the quantization noise
        self.last = {}

    def _noise(self, x):
        """model.py:48-49, same torch calls (so the CUDA generator stream matches the reference)."""
        if self.noise_override is not None:
            return self.noise_override
        q = torch.zeros(x.size(0), self.out_channel_N, x.size(2) // 16, x.size(3) // 16, device=x.device)
        return torch.nn.init.uniform_(torch.zeros_like(q), -0.5, 0.5)

    def forward(self, input_image):
        x = input_image
        if x.dim() != 4 or x.size(1) != 3:
            raise ValueError(f"expected [B,3,H,W] input, got {tuple(x.shape)}")
        if x.size(2) % 16 or x.size(3) % 16:
            # the reference fails here too (noise is H//16 but the convs give ceil(H/16), model.py:48,54)
            raise RuntimeError(f"H and W must be multiples of 16, got {x.size(2)}x{x.size(3)}")
        feature = self.Encoder(x)
        noise = self._noise(x) if self.training else None
        z, total_bits, image_bits, plane_bits = self.bitEstimator.bits(feature, noise)
        recon_image = self.Decoder(z)
        clipped_recon_image, sse_total, sse = B2F.distortion(recon_image, x)
        mse_loss = sse_total / x.numel()                                   # model.py:61
        bpp_feature = total_bits / (x.size(0) * x.size(2) * x.size(3))      # model.py:78
        self.last = {"image_bits": image_bits, "plane_bits": plane_bits, "sse": sse, "mse_loss": mse_loss}
        if self.return_mode == "mse":
            return clipped_recon_image, mse_loss, bpp_feature
        return clipped_recon_image, z, bpp_feature
