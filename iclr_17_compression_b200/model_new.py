"""Mirror of the reference's model_new.py: ImageCompressor_new (model_new.py:39-63) - the Balle-2017 transforms with a
sigmoid + straight-through binarizer instead of the noise / rounding quantizer and no rate term:

    code, pre_threshold = Encoder(x)            # Analysis_net_17_new: conv stack -> sigmoid -> (x > 0.5)
    z = code (training)  |  round(code) = code (eval)
    return clamp(Decoder(z), 0, 1), z, pre_threshold

eval + no_grad + out_channel_N in {128, 192, 256} runs the tcgen05 pipeline (bf16x3 analysis so the binary code is
exact away from the threshold, bf16 synthesis); everything else runs the exact-fp32 kernels with autograd."""
import torch
import torch.nn as nn

from . import functional as B2F
from . import tc as B2TC
from .model import *  # noqa: F401,F403  (star-import surface of model_new.py:1-16)
from .model import ImageCompressor, load_model, save_model  # noqa: F401
from .models import Analysis_net_17_new


class ImageCompressor_new(ImageCompressor):
    def _make_encoder(self, out_channel_N):
        return Analysis_net_17_new(out_channel_N=out_channel_N)

    def forward(self, input_image):
        x = input_image
        if x.dim() != 4 or x.size(1) != 3:
            raise ValueError(f"expected [B,3,H,W] input, got {tuple(x.shape)}")
        if x.size(2) % 16 or x.size(3) % 16:
            raise RuntimeError(f"H and W must be multiples of 16, got {x.size(2)}x{x.size(3)}")
        if self.training:
            self._maybe_join_process_group()       # multi-process data parallel under the launcher
        if self._use_tc(x):
            self.last_path = "tcgen05"
            x = B2F._prep(x.float(), "input image")
            y = self.encode_tc(x)                                    # conv3 output, fp32 NCHW
            code, pre_threshold, code_nhwc = B2F.sigmoid_binarize_infer(y)
            clipped, sse = self.decode_tc(code, ref_image=x, z_nhwc=code_nhwc)    # round(code) == code
            self.last = {"sse": sse}
            return clipped, code, pre_threshold
        self.last_path = "fp32"
        code, pre_threshold = self._enc(x)                            # model_new.py:48
        z = code                                                      # model_new.py:51-54: round() of 0 / 1 is a no-op
        recon = self._dec(z)
        clipped, _, sse = B2F.distortion(recon, x)                    # model_new.py:57 clamp(0, 1)
        self.last = {"sse": sse}
        return clipped, z, pre_threshold
