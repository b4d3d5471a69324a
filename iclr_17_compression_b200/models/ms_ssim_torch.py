"""Mirror of the reference's models/ms_ssim_torch.py (ssim, ms_ssim, SSIM, MS_SSIM; same signatures and defaults)
computed on the GPU by libb2c (b2c_ssim_level / b2c_avgpool2 / b2c_msssim_combine): one fused kernel per scale instead
of ten depth-wise convolutions and ~30 element-wise launches, and no device-to-host copy of the images
(train.py:178 calls ms_ssim(recon.cpu(), input.cpu(), data_range=1.0)).  CPU tensors are accepted for drop-in
compatibility - they are moved to the current CUDA device, the result comes back on the CPU; there is no CPU
implementation.  Forward only (the reference's validation loop runs under no_grad)."""
import ctypes

import torch

from .. import _lib
from ..functional import _p, _stream


def _fspecial_gauss_1d(size, sigma):
    """ms_ssim_torch.py:6-18"""
    coords = torch.arange(size).to(dtype=torch.float)
    coords -= size // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    g /= g.sum()
    return g.unsqueeze(0).unsqueeze(0)


def _check(X, Y, win_size):
    if len(X.shape) != 4:
        raise ValueError('Input images must 4-d tensor.')
    if not X.type() == Y.type():
        raise ValueError('Input images must have the same dtype.')
    if not X.shape == Y.shape:
        raise ValueError('Input images must have the same dimensions.')
    if not (win_size % 2 == 1):
        raise ValueError('Window size must be odd.')


def _to_device(X, Y):
    was_cpu = not X.is_cuda
    if was_cpu:
        if not torch.cuda.is_available():
            raise _lib.B2CError("ms_ssim runs on libb2c's CUDA kernels; no GPU is visible (this package has no CPU path)")
        X, Y = X.cuda(), Y.cuda()
    return X.detach().contiguous().float(), Y.detach().contiguous().float(), was_cpu


def _taps(win, win_size, win_sigma, dev):
    if win is None:
        return _fspecial_gauss_1d(win_size, win_sigma).reshape(-1).to(dev)
    return win.reshape(win.shape[0], -1)[0].detach().float().contiguous().to(dev)     # one 1-D kernel, repeated per channel


def _level_sums(X, Y, taps, data_range, sums_out):
    """sums_out[B,2] <- per-image sums of the ssim and cs maps; returns the number of map elements per image"""
    lib = _lib.load()
    B, C, H, W = X.shape
    ws = taps.numel()
    nbytes = lib.b2c_ssim_workspace_bytes(B, C, H, W, ws)
    if nbytes == 0:
        raise _lib.B2CError(f"image {H}x{W} is smaller than the {ws}-tap SSIM window")
    work = torch.empty((nbytes,), device=X.device, dtype=torch.uint8)
    C1, C2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    with torch.cuda.device(X.device):
        _lib.check(lib.b2c_ssim_level(_p(X), _p(Y), _p(taps), ws, ctypes.c_float(C1), ctypes.c_float(C2), _p(sums_out),
                                      B, C, H, W, _p(work), nbytes, _stream()), "ssim_level")
    return C * (H - ws + 1) * (W - ws + 1)


def _pool(X):
    B, C, H, W = X.shape
    ph, pw = H % 2, W % 2
    out = torch.empty((B, C, (H + 2 * ph - 2) // 2 + 1, (W + 2 * pw - 2) // 2 + 1), device=X.device, dtype=torch.float32)
    with torch.cuda.device(X.device):
        _lib.check(_lib.load().b2c_avgpool2(_p(X), _p(out), B * C, H, W, ph, pw, _stream()), "avgpool2")
    return out


def ssim(X, Y, win_size=11, win_sigma=1.5, win=None, data_range=255, size_average=True, full=False):
    """ms_ssim_torch.py:78-121"""
    _check(X, Y, win_size)
    X, Y, was_cpu = _to_device(X, Y)
    taps = _taps(win, win_size, win_sigma, X.device)
    sums = torch.empty((X.shape[0], 2), device=X.device, dtype=torch.float32)
    n = _level_sums(X, Y, taps, float(data_range), sums)
    ssim_val, cs = sums[:, 0] / n, sums[:, 1] / n
    if size_average:
        ssim_val, cs = ssim_val.mean(), cs.mean()
    if was_cpu:
        ssim_val, cs = ssim_val.cpu(), cs.cpu()
    return (ssim_val, cs) if full else ssim_val


def ms_ssim(X, Y, win_size=11, win_sigma=1.5, win=None, data_range=255, size_average=True, full=False, weights=None):
    """ms_ssim_torch.py:124-178.  Note the reference multiplies ssim_last^w_last into every one of the first L-1 rows
    before the product (ms_ssim_torch.py:169-170), i.e. that factor enters L-1 times; reproduced as written.  Unlike the
    reference (whose NaN check fails on a vector), size_average=False works for batches."""
    _check(X, Y, win_size)
    X, Y, was_cpu = _to_device(X, Y)
    dev = X.device
    if weights is None:
        weights = torch.tensor([0.0448, 0.2856, 0.3001, 0.2363, 0.1333], dtype=torch.float32)
    weights = weights.detach().float().contiguous().to(dev)
    taps = _taps(win, win_size, win_sigma, dev)
    levels, B = weights.shape[0], X.shape[0]
    sums = torch.empty((levels, B, 2), device=dev, dtype=torch.float32)
    counts = []
    for l in range(levels):
        counts.append(float(_level_sums(X, Y, taps, float(data_range), sums[l])))
        if l + 1 < levels:                     # the reference also pools after the last level; the result is unused
            X, Y = _pool(X), _pool(Y)
    counts = torch.tensor(counts, dtype=torch.float32).to(dev)
    out = torch.empty((B,), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().b2c_msssim_combine(_p(sums), _p(counts), _p(weights), _p(out), levels, B, _stream()),
                   "msssim_combine")
    if size_average:
        out = out.mean()
    return out.cpu() if was_cpu else out


class SSIM(torch.nn.Module):
    """ms_ssim_torch.py:181-198"""

    def __init__(self, win_size=11, win_sigma=1.5, data_range=None, size_average=True, channel=3):
        super().__init__()
        self.win = _fspecial_gauss_1d(win_size, win_sigma).repeat(channel, 1, 1, 1)
        self.size_average = size_average
        self.data_range = data_range

    def forward(self, X, Y):
        return ssim(X, Y, win=self.win, data_range=self.data_range, size_average=self.size_average)


class MS_SSIM(torch.nn.Module):
    """ms_ssim_torch.py:201-241"""

    def __init__(self, win_size=11, win_sigma=1.5, data_range=None, size_average=True, channel=3, weights=None):
        super().__init__()
        self.win = _fspecial_gauss_1d(win_size, win_sigma).repeat(channel, 1, 1, 1)
        self.size_average = size_average
        self.data_range = data_range
        self.weights = weights

    def forward(self, X, Y):
        return ms_ssim(X, Y, win=self.win, size_average=self.size_average, data_range=self.data_range,
                       weights=self.weights)
