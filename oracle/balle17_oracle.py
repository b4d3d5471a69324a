"""CPU oracle for the Balle-2017 codec hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-torch-on-CPU restatement of the reference algorithm.  It is
the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product package (``iclr_17_compression_b200``) never does; it fails loudly when
its CUDA library is missing.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the real
reference from /root/reference (with the two shims of SURVEY.md section 8c),
runs it on seeded inputs/weights and commits the outputs under ``tests/golden``;
``tests/test_oracle_golden.py`` asserts this restatement reproduces them
(bit-exact for the rounded latents, 1e-6 relative elsewhere).

Every function cites the reference file:line (paths relative to the reference
root) it follows.  Everything works on plain tensors held in a ``state_dict``
style mapping (same 30 keys as the reference), in fp32
(parity) or fp64 (tolerance budgeting: ``sd64 = to_dtype(sd, torch.float64)``).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

# models/GDN.py:32-50 constants
REPARAM_OFFSET = 2.0 ** -18
PEDESTAL = REPARAM_OFFSET ** 2
BETA_MIN = 1e-6
BETA_BOUND = (BETA_MIN + REPARAM_OFFSET ** 2) ** 0.5
GAMMA_BOUND = REPARAM_OFFSET


class _LowerBound(torch.autograd.Function):
    """models/GDN.py:10-24.  max(x, bound); grad passes iff x >= bound or g < 0."""

    @staticmethod
    def forward(ctx, inputs, bound):
        b = torch.ones_like(inputs) * bound
        ctx.save_for_backward(inputs, b)
        return torch.max(inputs, b)

    @staticmethod
    def backward(ctx, grad_output):
        inputs, b = ctx.saved_tensors
        keep = (inputs >= b) | (grad_output < 0)
        return keep.type(grad_output.dtype) * grad_output, None


def gdn_effective_params(beta_p: torch.Tensor, gamma_p: torch.Tensor):
    """models/GDN.py:73-79: bound + reparametrise beta/gamma."""
    beta = _LowerBound.apply(beta_p, BETA_BOUND) ** 2 - PEDESTAL
    gamma = _LowerBound.apply(gamma_p, GAMMA_BOUND) ** 2 - PEDESTAL
    return beta, gamma


def gdn(x: torch.Tensor, beta_p: torch.Tensor, gamma_p: torch.Tensor, inverse: bool) -> torch.Tensor:
    """models/GDN.py:64-94.  norm_i = sqrt(beta_i + sum_j gamma[i,j] x_j^2);
    y = x*norm (inverse) or x/norm."""
    ch = x.shape[1]
    beta, gamma = gdn_effective_params(beta_p, gamma_p)
    norm = torch.sqrt(F.conv2d(x ** 2, gamma.view(ch, ch, 1, 1), beta))
    return x * norm if inverse else x / norm


def analysis(x: torch.Tensor, sd: Dict[str, torch.Tensor], prefix: str = "Encoder.") -> torch.Tensor:
    """models/analysis_17.py:32-39 (ctor geometry :14,18,22)."""
    p = prefix
    y = F.conv2d(x, sd[p + "conv1.weight"], sd[p + "conv1.bias"], stride=4, padding=4)
    y = gdn(y, sd[p + "gdn1.beta"], sd[p + "gdn1.gamma"], False)
    y = F.conv2d(y, sd[p + "conv2.weight"], sd[p + "conv2.bias"], stride=2, padding=2)
    y = gdn(y, sd[p + "gdn2.beta"], sd[p + "gdn2.gamma"], False)
    y = F.conv2d(y, sd[p + "conv3.weight"], None, stride=2, padding=2)
    return y


def synthesis(z: torch.Tensor, sd: Dict[str, torch.Tensor], prefix: str = "Decoder.") -> torch.Tensor:
    """models/synthesis_17.py:27-31 (ctor geometry :15,19,23)."""
    p = prefix
    r = F.conv_transpose2d(z, sd[p + "deconv1.weight"], sd[p + "deconv1.bias"], stride=2, padding=2, output_padding=1)
    r = gdn(r, sd[p + "igdn1.beta"], sd[p + "igdn1.gamma"], True)
    r = F.conv_transpose2d(r, sd[p + "deconv2.weight"], sd[p + "deconv2.bias"], stride=2, padding=2, output_padding=1)
    r = gdn(r, sd[p + "igdn2.beta"], sd[p + "igdn2.gamma"], True)
    r = F.conv_transpose2d(r, sd[p + "deconv3.weight"], sd[p + "deconv3.bias"], stride=4, padding=4, output_padding=3)
    return r


def bitparm(x, h, b, a=None):
    """models/bitEstimator.py:20-25."""
    if a is None:
        return torch.sigmoid(x * F.softplus(h) + b)
    x = x * F.softplus(h) + b
    return x + torch.tanh(x) * torch.tanh(a)


def bit_estimator(x: torch.Tensor, sd: Dict[str, torch.Tensor], prefix: str = "bitEstimator.") -> torch.Tensor:
    """models/bitEstimator.py:38-42: CDF = f4(f3(f2(f1(x))))."""
    p = prefix
    for k in (1, 2, 3):
        x = bitparm(x, sd[f"{p}f{k}.h"], sd[f"{p}f{k}.b"], sd[f"{p}f{k}.a"])
    return bitparm(x, sd[p + "f4.h"], sd[p + "f4.b"], None)


def estimate_bits(z: torch.Tensor, sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """model.py:71-74 (iclr18_estimate_bits_z)."""
    prob = bit_estimator(z + 0.5, sd) - bit_estimator(z - 0.5, sd)
    return torch.sum(torch.clamp(-1.0 * torch.log(prob + 1e-10) / math.log(2.0), 0, 50))


def quantize(y: torch.Tensor, training: bool, noise: Optional[torch.Tensor]) -> torch.Tensor:
    """model.py:53-56: additive U(-.5,.5) noise (training) / round-half-even (eval)."""
    if training:
        assert noise is not None
        return y + noise
    return torch.round(y)


def forward(x: torch.Tensor, sd: Dict[str, torch.Tensor], training: bool = False,
            noise: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """model.py:47-80.  Returns every intermediate the parity tests look at.

    ``clipped, z, bpp`` is the HEAD return tuple (model.py:80); ``mse`` is the
    value computed at model.py:61 on the UNclamped reconstruction (the value
    the commented-out return at model.py:81 hands to train.py:97)."""
    y = analysis(x, sd)
    z = quantize(y, training, noise)
    recon = synthesis(z, sd)
    clipped = recon.clamp(0.0, 1.0)
    mse = torch.mean((recon - x).pow(2))
    bits = estimate_bits(z, sd)
    bpp = bits / (x.shape[0] * x.shape[2] * x.shape[3])
    return {"y": y, "z": z, "recon": recon, "clipped": clipped, "mse": mse, "bits": bits, "bpp": bpp}


class _Binarizer(torch.autograd.Function):
    """models/binarizer.py:7-27: (i > 0.5).float(), straight-through gradient."""

    @staticmethod
    def forward(ctx, i):
        return (i > 0.5).to(i.dtype)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output


def forward_new(x: torch.Tensor, sd: Dict[str, torch.Tensor], training: bool = False) -> Dict[str, torch.Tensor]:
    """model_new.py:47-61 (ImageCompressor_new) over models/analysis_17_new.py:32-38: the analysis transform, sigmoid,
    binarizer; z = code (training) / round(code) (eval); synthesis; clamp.  Returns the HEAD tuple members
    (clipped, z, pre_threshold) plus the intermediates."""
    y = analysis(x, sd)
    pre = torch.sigmoid(y)
    code = _Binarizer.apply(pre)
    z = code if training else torch.round(code)
    recon = synthesis(z, sd)
    return {"y": y, "pre_threshold": pre, "z": z, "recon": recon, "clipped": recon.clamp(0.0, 1.0)}


def new_loss_and_grads(x, sd, target):
    """a loss touching all three outputs of ImageCompressor_new the way train_new.py:161-180 combines them (image MSE
    + code MSE against another code) : mean((clipped - x)^2) + 0.5 mean((z - target)^2) + 0.1 mean(pre_threshold)"""
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items() if not k.startswith("bitEstimator.")}
    out = forward_new(x, leaf, training=True)
    loss = torch.mean((out["clipped"] - x) ** 2) + 0.5 * torch.mean((out["z"] - target) ** 2) + 0.1 * out["pre_threshold"].mean()
    keys = list(leaf.keys())
    grads = torch.autograd.grad(loss, [leaf[k] for k in keys])
    return out, loss.detach(), dict(zip(keys, grads))


def rd_loss_and_grads(x, sd, noise, lam: float = 8192.0):
    """train.py:102-105: rd_loss = lambda*mse + bpp; grads for all 30 tensors."""
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    out = forward(x, leaf, training=True, noise=noise)
    loss = lam * out["mse"] + out["bpp"]
    keys = list(leaf.keys())
    grads = torch.autograd.grad(loss, [leaf[k] for k in keys])
    return out, loss.detach(), dict(zip(keys, grads))


# --------------------------------------------------------------------------------------
# SSIM / MS-SSIM (models/ms_ssim_torch.py), used by the validation loop train.py:178
# --------------------------------------------------------------------------------------
def _gauss_1d(size: int, sigma: float) -> torch.Tensor:
    """models/ms_ssim_torch.py:6-18."""
    coords = torch.arange(size).to(dtype=torch.float)
    coords -= size // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def _gaussian_filter(x: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    """models/ms_ssim_torch.py:21-35: depth-wise 1 x k then k x 1 convolution, valid padding."""
    C = x.shape[1]
    w = g.to(x.dtype).view(1, 1, 1, -1).repeat(C, 1, 1, 1)
    out = F.conv2d(x, w, stride=1, padding=0, groups=C)
    return F.conv2d(out, w.transpose(2, 3), stride=1, padding=0, groups=C)


def ssim_per_image(X, Y, g, data_range):
    """models/ms_ssim_torch.py:38-75 with size_average=False, full=True: (ssim[B], cs[B])."""
    C1, C2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    mu1, mu2 = _gaussian_filter(X, g), _gaussian_filter(Y, g)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    sigma1_sq = _gaussian_filter(X * X, g) - mu1_sq
    sigma2_sq = _gaussian_filter(Y * Y, g) - mu2_sq
    sigma12 = _gaussian_filter(X * Y, g) - mu1_mu2
    cs_map = (2 * sigma12 + C2) / (sigma1_sq + sigma2_sq + C2)
    ssim_map = ((2 * mu1_mu2 + C1) / (mu1_sq + mu2_sq + C1)) * cs_map
    return ssim_map.mean(-1).mean(-1).mean(-1), cs_map.mean(-1).mean(-1).mean(-1)


def ms_ssim_per_image(X, Y, data_range=1.0, win_size=11, win_sigma=1.5,
                      weights=(0.0448, 0.2856, 0.3001, 0.2363, 0.1333)):
    """models/ms_ssim_torch.py:124-178 with size_average=False: ms-ssim[B]."""
    g = _gauss_1d(win_size, win_sigma)
    w = torch.tensor(weights, dtype=X.dtype)
    mcs = []
    for _ in range(len(weights)):
        ssim_val, cs = ssim_per_image(X, Y, g, data_range)
        mcs.append(cs)
        padding = (X.shape[2] % 2, X.shape[3] % 2)
        X = F.avg_pool2d(X, kernel_size=2, padding=padding)
        Y = F.avg_pool2d(Y, kernel_size=2, padding=padding)
    mcs = torch.stack(mcs, dim=0)
    return torch.prod((mcs[:-1] ** w[:-1].unsqueeze(1)) * (ssim_val ** w[-1]), dim=0)


def synth_pair(B, H, W, seed, noise=0.05):
    """an image and a degraded copy for the SSIM tests: smooth image + clamped Gaussian noise"""
    Hp, Wp = (H + 7) // 8 * 8, (W + 7) // 8 * 8
    x = synth_images("S", B, Hp, Wp, seed)[:, :, :H, :W].contiguous()
    g = torch.Generator().manual_seed(seed + 17)
    y = (x + noise * torch.randn(x.shape, generator=g)).clamp(0, 1)
    return x, y


# --------------------------------------------------------------------------------------
# weights: the reference's constructor order and init (so the RNG stream matches)
# --------------------------------------------------------------------------------------
def init_state_dict(out_channel_N: int = 128, seed: Optional[int] = 234) -> Dict[str, torch.Tensor]:
    """Seed-for-seed reproduction of ``ImageCompressor(out_channel_N)`` initial weights.

    Follows model.py:41-43 (Encoder, Decoder, bitEstimator in that order),
    models/analysis_17.py:14-23, models/synthesis_17.py:15-25,
    models/GDN.py:46-62, models/bitEstimator.py:13-16.  Real torch modules are
    constructed in the same order because their default (kaiming) init consumes
    the global generator before the xavier re-init does."""
    import torch.nn as nn
    N = out_channel_N
    if seed is not None:
        torch.manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def gdn_params(name):
        sd[name + ".beta"] = torch.sqrt(torch.ones(N) + PEDESTAL)
        sd[name + ".gamma"] = torch.sqrt(0.1 * torch.eye(N) + PEDESTAL)

    def conv(name, cin, cout, k, s, p, gain, bias=True, transposed=False, op=0):
        m = (nn.ConvTranspose2d(cin, cout, k, stride=s, padding=p, output_padding=op) if transposed
             else nn.Conv2d(cin, cout, k, stride=s, padding=p, bias=bias))
        torch.nn.init.xavier_normal_(m.weight.data, gain)
        sd[name + ".weight"] = m.weight.data
        if bias:
            torch.nn.init.constant_(m.bias.data, 0.01)
            sd[name + ".bias"] = m.bias.data

    conv("Encoder.conv1", 3, N, 9, 4, 4, math.sqrt(2 * (3 + N) / 6))
    gdn_params("Encoder.gdn1")
    conv("Encoder.conv2", N, N, 5, 2, 2, math.sqrt(2))
    gdn_params("Encoder.gdn2")
    conv("Encoder.conv3", N, N, 5, 2, 2, math.sqrt(2), bias=False)
    conv("Decoder.deconv1", N, N, 5, 2, 2, math.sqrt(2), transposed=True, op=1)
    gdn_params("Decoder.igdn1")
    conv("Decoder.deconv2", N, N, 5, 2, 2, math.sqrt(2), transposed=True, op=1)
    gdn_params("Decoder.igdn2")
    conv("Decoder.deconv3", N, 3, 9, 4, 4, math.sqrt(2), transposed=True, op=3)
    for k in (1, 2, 3, 4):
        sd[f"bitEstimator.f{k}.h"] = torch.nn.init.normal_(torch.empty(N).view(1, -1, 1, 1), 0, 0.01)
        sd[f"bitEstimator.f{k}.b"] = torch.nn.init.normal_(torch.empty(N).view(1, -1, 1, 1), 0, 0.01)
        if k != 4:
            sd[f"bitEstimator.f{k}.a"] = torch.nn.init.normal_(torch.empty(N).view(1, -1, 1, 1), 0, 0.01)
    return sd


def perturb_state_dict(sd: Dict[str, torch.Tensor], seed: int = 3003) -> Dict[str, torch.Tensor]:
    """W1 weights: W0 deterministically perturbed so every term is
    exercised - dense gamma, non-unit beta, a non-flat CDF, and a few beta/gamma
    entries below their LowerBound.  Not part of the reference; shared by the
    golden generator and the tests so both see identical weights."""
    g = torch.Generator().manual_seed(seed)
    out = {k: v.clone() for k, v in sd.items()}
    for k in sorted(out.keys()):
        v = out[k]
        if k.endswith(".gamma"):
            eff = v ** 2 - PEDESTAL + 0.05 * torch.randn(v.shape, generator=g).abs()
            v = torch.sqrt(eff + PEDESTAL)
            idx = torch.randint(0, v.numel(), (7,), generator=g)
            v.view(-1)[idx] = GAMMA_BOUND * 0.25          # below the bound -> LowerBound active
            out[k] = v
        elif k.endswith(".beta"):
            eff = (v ** 2 - PEDESTAL) * (0.5 + 1.5 * torch.rand(v.shape, generator=g))
            v = torch.sqrt(eff + PEDESTAL)
            idx = torch.randint(0, v.numel(), (3,), generator=g)
            v[idx] = BETA_BOUND * 0.5
            out[k] = v
        elif k.startswith("bitEstimator."):
            std = 1.0 if k.endswith(".h") else 0.5
            out[k] = torch.randn(v.shape, generator=g) * std
        elif k.endswith(".bias"):
            out[k] = v + 0.05 * torch.randn(v.shape, generator=g)
    return out


def to_dtype(sd, dtype):
    return {k: v.to(dtype) for k, v in sd.items()}


# --------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------
def synth_images(kind: str, B: int, H: int, W: int, seed: int) -> torch.Tensor:
    """'U' = iid uniform [0,1); 'S' = smooth (bicubic x8 upsample of uniform noise, clamped)."""
    g = torch.Generator().manual_seed(seed)
    if kind == "U":
        return torch.rand(B, 3, H, W, generator=g)
    if kind == "S":
        low = torch.rand(B, 3, H // 8, W // 8, generator=g)
        return F.interpolate(low, scale_factor=8, mode="bicubic", align_corners=False).clamp(0, 1).contiguous()
    raise ValueError(kind)


def synth_noise(B: int, N: int, H: int, W: int, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, N, H // 16, W // 16, generator=g) - 0.5
