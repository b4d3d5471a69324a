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

from . import _lib
from . import functional as B2F
from . import tc as B2TC
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

    # attribute (and state_dict key) names of the two transforms: model_twoSteps.py:41-42 calls them Encoder1 / Decoder1
    _ENC, _DEC = "Encoder", "Decoder"

    def _make_encoder(self, out_channel_N):
        return Analysis_net_17(out_channel_N=out_channel_N)

    @property
    def _enc(self):
        return getattr(self, self._ENC)

    @property
    def _dec(self):
        return getattr(self, self._DEC)

    def canonical_named_parameters(self):
        """named_parameters() with the transforms' prefixes normalised to Encoder. / Decoder. (the names the
        tensor-core training step is written against)"""
        out = []
        for k, p in self.named_parameters():
            if k.startswith(self._ENC + "."):
                k = "Encoder." + k[len(self._ENC) + 1:]
            elif k.startswith(self._DEC + "."):
                k = "Decoder." + k[len(self._DEC) + 1:]
            out.append((k, p))
        return out

    def __init__(self, out_channel_N=128, return_mode=None):
        super().__init__()
        setattr(self, self._ENC, self._make_encoder(out_channel_N))
        setattr(self, self._DEC, Synthesis_net_17(out_channel_N=out_channel_N))
        self.bitEstimator = BitEstimator(channel=out_channel_N)
        self.out_channel_N = out_channel_N
        mode = return_mode or os.environ.get("IMAGECOMPRESSOR_RETURN", "latent")
        if mode not in ("latent", "mse"):
            raise ValueError("return_mode must be 'latent' (model.py:80) or 'mse' (model.py:81)")
        self.return_mode = mode
        self.noise_override = None      # test hook: inject the quantization noise
        self.last = {}
        # eval + no_grad + out_channel_N in {128, 192, 256} runs the tcgen05 pipeline (bf16x3 analysis, bf16
        # synthesis); everything else (autograd in eval mode, other widths) runs the exact-fp32 kernels.
        self.fast_inference = True
        # training + out_channel_N == 128: bf16 tensor-core forward/backward (train_tc.TrainStepTC); gradients
        # match the fp32 path within 1e-2 relative L2.  Set False for the exact-fp32 kernels.
        self.fast_training = os.environ.get("B2C_FAST_TRAINING", "1") == "1"
        # ... and, from the second step with the same input shape on, as two CUDA-graph replays (forward, backward)
        # under autograd: the reference's own loop (train.py:97-112) is host bound otherwise (train_tc._StepGraphs).
        # One set of saved activations per shape: backward() must run before the next training forward.
        self.graph_training = os.environ.get("B2C_TRAIN_AUTOGRAPH", "1") == "1"
        # eval + no_grad on small batches: the tensor-core pipeline as ONE graph replay (_forward_tc_maybe_graphed)
        self.graph_inference = os.environ.get("B2C_EVAL_AUTOGRAPH", "1") == "1"
        self.last_path = None
        self._packed = None
        self._packed_key = None
        self._grad_sync = None
        self._zflag = None              # functional.RangeFlag of the tensor-core forward (|z| > 256 guard)

    def _maybe_join_process_group(self):
        """multi-process data parallel without touching train.py: when the launcher
        (run_reference_script.py) set B2C_DDP=1, join the NCCL group on the first training forward,
        take rank 0's weights and all-reduce gradients from backward hooks."""
        if self._grad_sync is not None or os.environ.get("B2C_DDP") != "1":
            return
        if int(os.environ.get("WORLD_SIZE", "1")) <= 1:
            return
        import torch.distributed as dist
        from .ddp import GradSync
        if not dist.is_initialized():
            dist.init_process_group("nccl" if next(self.parameters()).is_cuda else "gloo")
        self._grad_sync = GradSync(self, auto_finish=True)

    # ------------------------------------------------------------------ tensor-core inference path
    def _pack_key(self):
        # _version does not move when the fused optimizer writes through raw pointers: weights_epoch() does
        return (_lib.weights_epoch(),) + tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _packed_weights(self):
        """bf16 (hi, lo) tap-major weight panels and effective GDN parameters, re-packed whenever a
        parameter tensor is replaced or modified in place (keyed by data_ptr and _version, so
        nn.DataParallel replicas and optimizer steps are handled)."""
        key = self._pack_key()
        if self._packed is None or key != self._packed_key:
            E, D = self._enc, self._dec
            self._packed = {
                "conv1": B2TC.pack_weights(B2TC.TC_CONV1, E.conv1.weight, True),
                "conv2": B2TC.pack_weights(B2TC.TC_CONV5S2, E.conv2.weight, True),
                "conv3": B2TC.pack_weights(B2TC.TC_CONV5S2, E.conv3.weight, True),
                "gdn1": B2TC.pack_gdn(E.gdn1.beta, E.gdn1.gamma, True),
                "gdn2": B2TC.pack_gdn(E.gdn2.beta, E.gdn2.gamma, True),
                "deconv1": B2TC.pack_weights(B2TC.TC_DECONV5S2, D.deconv1.weight, False),
                "deconv2": B2TC.pack_weights(B2TC.TC_DECONV5S2, D.deconv2.weight, False),
                "deconv3": B2TC.pack_weights(B2TC.deconv3_kind(self.out_channel_N), D.deconv3.weight, False),
                "igdn1": B2TC.pack_gdn(D.igdn1.beta, D.igdn1.gamma, False),
                "igdn2": B2TC.pack_gdn(D.igdn2.beta, D.igdn2.gamma, False),
                # bits[c][z] for integer z: the inference entropy pass becomes a table read (functional.entropy_lut)
                "bits_lut": B2F.entropy_lut([p.detach().reshape(-1) for p in self.bitEstimator.param_list()]),
            }
            self._packed_key = key
        return self._packed

    def encode_tc(self, x, latent=None, reduce_bits=True):
        """analysis transform on tensor cores (error-compensated bf16x3) -> y, fp32 NCHW.  ``latent`` (see
        tc.layer_run): quantizer + entropy model in the last layer's epilogue -> (z, z bf16 NHWC, image_bits)."""
        P, E = self._packed_weights(), self._enc
        h = B2TC.layer_run(B2TC.TC_CONV1, x, None, P["conv1"], E.conv1.bias, gdn=P["gdn1"], gdn_mode=1, split3=True)
        h = B2TC.layer_run(B2TC.TC_CONV5S2, h[0], h[1], P["conv2"], E.conv2.bias, gdn=P["gdn2"], gdn_mode=1,
                           split3=True)
        return B2TC.layer_run(B2TC.TC_CONV5S2, h[0], h[1], P["conv3"], None, split3=True, out_f32_nchw=True,
                              latent=latent, reduce_sse=reduce_bits)

    def check_latent_range(self):
        """The tensor-core synthesis transform reads the rounded latent as bf16, exact only for |z| <= 256.  The
        inference quantizer raises a sticky flag in mapped pinned memory when a latent exceeds that; it is polled
        here without synchronising - at the start of every tensor-core forward and wherever a caller has synchronised
        anyway (HostPipeline.compress).  On overflow the model switches itself to the exact-fp32 kernels and raises,
        so a wrong reconstruction cannot pass silently."""
        if self._zflag is not None and self._zflag.raised():
            self._zflag.clear()
            self.fast_inference = False
            raise _lib.B2CError("a tensor-core forward of this model produced latents beyond +-256, which the bf16 "
                                "synthesis path cannot represent exactly: its reconstruction is invalid.  "
                                "fast_inference is now off (exact-fp32 kernels); run the batch again.")

    def decode_tc(self, z, ref_image=None, z_nhwc=None, reduce_sse=True):
        """synthesis transform on tensor cores (bf16) -> unclamped reconstruction, fp32 NCHW; with ref_image
        the distortion is fused into the last layer: returns (clamped reconstruction, sse[B,2]).
        A caller-supplied ``z`` without its bf16 copy is range-checked (one host sync): beyond +-256 the exact-fp32
        synthesis transform runs instead."""
        P, D = self._packed_weights(), self._dec
        if z_nhwc is None and float(z.abs().max()) > 256.0:
            recon = self._dec(z)
            if ref_image is None:
                return recon
            clipped, _, sse = B2F.distortion(recon, ref_image)
            return clipped, sse
        zh = z_nhwc if z_nhwc is not None else B2TC.nchw_to_nhwc(z, False)[0]   # integer |z| <= 256: exact in bf16
        h = B2TC.layer_run(B2TC.TC_DECONV5S2, zh, None, P["deconv1"], D.deconv1.bias, gdn=P["igdn1"], gdn_mode=2)
        h = B2TC.layer_run(B2TC.TC_DECONV5S2, h[0], None, P["deconv2"], D.deconv2.bias, gdn=P["igdn2"], gdn_mode=2)
        return B2TC.layer_run(B2TC.deconv3_kind(self.out_channel_N), h[0], None, P["deconv3"], D.deconv3.bias,
                              out_f32_nchw=True, ref_image=ref_image, reduce_sse=reduce_sse)

    def _forward_tc(self, x):
        """the tensor-core eval pipeline: 7 fused launches + 1 reduction"""
        bitp = [p.detach().reshape(-1) for p in self.bitEstimator.param_list()]
        lut = self._packed_weights()["bits_lut"]
        if self.out_channel_N == 128 and os.environ.get("B2C_FUSED_LATENT", "1") == "1" \
                and os.environ.get("B2C_TC_PERSIST", "1") != "0":
            # the quantizer and the entropy model run in the epilogue of the last analysis layer: no separate pass over y
            z, zh, image_bits, bits_partial = self.encode_tc(
                x, latent=(lut, B2F.LUT_HALF_RANGE, [B2F._prep(p, "bit-estimator parameter") for p in bitp], self._zflag),
                reduce_bits=False)                 # (the statistics launch at the end reduces the per-warp bit sums)
        else:
            bits_partial = None
            feature = self.encode_tc(x)
            z, zh, image_bits = B2F.quant_entropy_infer(feature, bitp, lut=lut, range_flag=self._zflag)
        # (reduce_sse=False: the per-tile partial sums come back and ONE launch forms the per-image squared errors, the
        # batch totals, mse and bpp - model.py:61,73,78)
        clipped, partial = self.decode_tc(z, ref_image=x, z_nhwc=zh, reduce_sse=False)
        sse, totals = B2F.infer_stats(partial, image_bits, x.numel(), x.size(0) * x.size(2) * x.size(3), bits_partial)
        return clipped, z, image_bits, sse, totals[0], totals[1], totals[2], totals[3]

    # Small eval batches (the reference's testKodak feeds ONE image per call, train.py:162-171) are host bound: the 8
    # launches with their TMA descriptors and output allocations cost ~0.4 ms against ~0.06 ms of device time per
    # Kodak image.  The second time an (input shape, weight version) pair is seen the pipeline is captured as a CUDA
    # graph; later calls copy the input into the graph's buffer, replay, and hand out copies of the results.
    _GRAPH_MAX_PIXELS = 2 * 1024 * 1024           # larger batches are device bound: not worth a resident graph

    def _forward_tc_maybe_graphed(self, x):
        if (not self.graph_inference or torch.cuda.is_current_stream_capturing()
                or x.size(0) * x.size(2) * x.size(3) > self._GRAPH_MAX_PIXELS):
            return self._forward_tc(x)
        cache = self.__dict__.setdefault("_eval_graphs", {})
        key = (tuple(x.shape), x.device.index, self._pack_key())
        hit = cache.get(key)
        if hit is None or hit == "failed":
            if hit is None:
                if len(cache) >= 4:
                    cache.pop(next(iter(cache)))
                cache[key] = 1
            return self._forward_tc(x)
        if hit == 1:
            try:
                self._packed_weights()                              # packing happens outside the graph
                xs = torch.empty_like(x)
                cur = torch.cuda.current_stream(x.device)
                side = torch.cuda.Stream(device=x.device)
                side.wait_stream(cur)
                with torch.cuda.stream(side):                       # warm-up on a side stream, as capture requires
                    xs.copy_(x)
                    self._forward_tc(xs)
                cur.wait_stream(side)
                torch.cuda.synchronize(x.device)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    outs = self._forward_tc(xs)
                hit = cache[key] = (g, xs, outs)
            except Exception as e:                                  # capture is an optimisation only
                import warnings
                warnings.warn(f"CUDA-graph capture of the eval forward failed ({e!r}); running it eagerly")
                cache[key] = "failed"
                return self._forward_tc(x)
        g, xs, outs = hit
        xs.copy_(x)
        g.replay()
        self.last_path = "tcgen05-graphed"
        return tuple(t.clone() for t in outs)      # the graph's buffers are overwritten by the next replay

    def _use_tc(self, x):
        return (self.fast_inference and not self.training and not torch.is_grad_enabled()
                and self.out_channel_N in (128, 192, 256) and x.is_cuda)

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
        if self.training:
            self._maybe_join_process_group()
        if (self.training and self.fast_training and self.out_channel_N == 128 and x.is_cuda
                and torch.is_grad_enabled()):
            from .train_tc import train_forward
            x = B2F._prep(x, "input image")
            clipped_recon_image, mse_loss, bpp_feature, z = train_forward(self, x, self._noise(x))
            self.last_path = "tcgen05-bf16-train" + ("-graphed" if self.__dict__.get("_last_train_graphed") else "")
            self.last = {"mse_loss": mse_loss.detach()}      # detached: must not keep the autograd graph alive
            if self.return_mode == "mse":
                return clipped_recon_image, mse_loss, bpp_feature
            return clipped_recon_image, z, bpp_feature
        tc_path = self._use_tc(x)
        self.last_path = "tcgen05" if tc_path else "fp32"
        if tc_path:
            self.check_latent_range()
            if self._zflag is None:
                self._zflag = B2F.RangeFlag()
            x = B2F._prep(x.float(), "input image")
            clipped_recon_image, z, image_bits, sse, total_bits, sse_total, mse_loss, bpp_feature = \
                self._forward_tc_maybe_graphed(x)
            self.last = {"image_bits": image_bits, "plane_bits": None, "sse": sse, "mse_loss": mse_loss.detach()}
            if self.return_mode == "mse":
                return clipped_recon_image, mse_loss, bpp_feature
            return clipped_recon_image, z, bpp_feature
        else:
            feature = self._enc(x)
            noise = self._noise(x) if self.training else None
            z, total_bits, image_bits, plane_bits = self.bitEstimator.bits(feature, noise)
            recon_image = self._dec(z)
            clipped_recon_image, sse_total, sse = B2F.distortion(recon_image, x)
        mse_loss = sse_total / x.numel()                                   # model.py:61
        bpp_feature = total_bits / (x.size(0) * x.size(2) * x.size(3))      # model.py:78
        self.last = {"image_bits": image_bits, "plane_bits": plane_bits, "sse": sse, "mse_loss": mse_loss.detach()}
        if self.return_mode == "mse":
            return clipped_recon_image, mse_loss, bpp_feature
        return clipped_recon_image, z, bpp_feature
