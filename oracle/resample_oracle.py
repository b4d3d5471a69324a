"""TEST INFRASTRUCTURE - CPU restatement of the training input transform of the reference:

    transforms.Compose([RandomResizedCrop(image_size), RandomHorizontalFlip(), RandomVerticalFlip(), ToTensor()])
                                                                                     (datasets.py:24-34 of the reference)

The arithmetic lives in two third-party dependencies the reference does not pin (no requirements file): torchvision
(`transforms.RandomResizedCrop.get_params`, `functional.resized_crop` = `img.crop(box)` then `img.resize(size,
BILINEAR)`, `hflip` / `vflip`, `to_tensor` = uint8 / 255) and Pillow (`Image.resize` -> `ImagingResample`,
src/libImaging/Resample.c: separable antialiased resampling of 8-bit channels in 22-bit fixed point, horizontal pass
first, each pass rounded and clipped to uint8).  This file restates their published algorithms in numpy / plain
Python; it is PINNED against the real packages of this image (Pillow 12.2.0, torchvision 0.26.0) through
tests/golden/augment_ref.npz (made by tests/golden/make_augment_golden.py) and, when they are importable, directly.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
import math

import numpy as np
import torch

PRECISION_BITS = 32 - 8 - 2          # Resample.c: coefficients are ints with 22 fractional bits


def bilinear_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter (support 1) over the whole input
    [0, in_size): -> (ksize, bounds int32 [out, 2] = (first input index, tap count), coeffs int32 [out, ksize])"""
    scale = in_size / out_size
    filterscale = scale if scale >= 1.0 else 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = []
        ww = 0.0
        for x in range(xmax):
            v = (x + xmin - center + 0.5) * ss
            if v < 0.0:
                v = -v
            wv = 1.0 - v if v < 1.0 else 0.0
            w.append(wv)
            ww += wv
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _pass(img, bounds, kk, axis):
    """one resampling pass over `axis` of a uint8 [H, W, C] image (ImagingResampleHorizontal/Vertical_8bpc)"""
    img = np.moveaxis(img, axis, 0).astype(np.int64)              # [in, other, C]
    out = np.empty((bounds.shape[0],) + img.shape[1:], np.uint8)
    for o in range(bounds.shape[0]):
        x0, n = int(bounds[o, 0]), int(bounds[o, 1])
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(kk[o, :n].astype(np.int64), img[x0:x0 + n], axes=(0, 0))
        out[o] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return np.moveaxis(out, 0, axis)


def resize_bilinear_u8(img, out_h, out_w):
    """PIL `Image.resize((out_w, out_h), BILINEAR)` of a uint8 [H, W, C] image (Resample.c ImagingResample: a pass is
    skipped when that dimension does not change; horizontal first)"""
    H, W = img.shape[:2]
    if W != out_w:
        _, b, k = bilinear_coeffs(W, out_w)
        img = _pass(img, b, k, 1)
    if H != out_h:
        _, b, k = bilinear_coeffs(H, out_h)
        img = _pass(img, b, k, 0)
    return img


def sample_params(height, width, scale=(0.08, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0)):
    """torchvision RandomResizedCrop.get_params followed by the two flip draws, on torch's global generator in the
    same order as the reference's Compose (datasets.py:27-30): -> (top, left, h, w, hflip, vflip)"""
    area = height * width
    log_ratio = torch.log(torch.tensor(ratio))
    box = None
    for _ in range(10):
        target_area = area * torch.empty(1).uniform_(scale[0], scale[1]).item()
        aspect_ratio = torch.exp(torch.empty(1).uniform_(log_ratio[0], log_ratio[1])).item()
        w = int(round(math.sqrt(target_area * aspect_ratio)))
        h = int(round(math.sqrt(target_area / aspect_ratio)))
        if 0 < w <= width and 0 < h <= height:
            i = torch.randint(0, height - h + 1, size=(1,)).item()
            j = torch.randint(0, width - w + 1, size=(1,)).item()
            box = (i, j, h, w)
            break
    if box is None:                                               # fallback to a central crop
        in_ratio = float(width) / float(height)
        if in_ratio < min(ratio):
            w = width
            h = int(round(w / min(ratio)))
        elif in_ratio > max(ratio):
            h = height
            w = int(round(h * max(ratio)))
        else:
            w, h = width, height
        box = ((height - h) // 2, (width - w) // 2, h, w)
    hflip = bool(torch.rand(1) < 0.5)
    vflip = bool(torch.rand(1) < 0.5)
    return box + (hflip, vflip)


def resized_crop_flip(img, top, left, h, w, hflip, vflip, size=256):
    """uint8 [H, W, 3] -> float32 [3, size, size] in [0, 1]: crop, PIL bilinear resize, flips, ToTensor"""
    out = resize_bilinear_u8(np.ascontiguousarray(img[top:top + h, left:left + w]), size, size)
    if hflip:
        out = out[:, ::-1]
    if vflip:
        out = out[::-1]
    return (torch.from_numpy(np.ascontiguousarray(out)).permute(2, 0, 1).contiguous().to(torch.float32).div(255))


def synth_photo(h, w, seed):
    """deterministic uint8 test image with smooth structure, edges and noise"""
    g = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.empty((h, w, 3), np.float64)
    for c in range(3):
        fx, fy, ph = g.uniform(0.01, 0.2, 3)
        img[..., c] = 127 + 90 * np.sin(xx * fx + ph) * np.cos(yy * fy) + 60 * ((xx // 17 + yy // 23) % 2) - 30
    img += g.normal(0, 12, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)
