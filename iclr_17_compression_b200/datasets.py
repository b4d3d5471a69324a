"""Mirror of the reference's ``datasets.py`` (``Datasets`` :14-37, ``get_loader`` :40-51, ``get_train_loader`` :54-61,
``TestKodakDataset`` :63-80) with the per-image CPU transform moved onto the GPU.

The reference runs, per image and on one CPU worker, ``Compose([RandomResizedCrop(image_size),
RandomHorizontalFlip(), RandomVerticalFlip(), ToTensor()])`` (``datasets.py:27-32``).  Here ``__getitem__`` only decodes
the file (PIL, host) and draws the crop box and flips with the SAME calls on torch's global generator, in the same
order, as torchvision does; a batch of decoded uint8 images then goes to the device in one pinned copy and ONE launch of
``b2c_resized_crop_flip`` (``csrc/b2c_augment.cu``) produces the fp32 ``[B, 3, S, S]`` batch, bit-identical to the
reference transform (Pillow's antialiased bilinear resample in 22-bit fixed point, flips, ``/ 255``).  CUDA only.

    train_dataset, train_loader = get_train_loader(data_dir, 256, batch_size)     # as in train.py
    for x in train_loader:            # x: CUDA fp32 [B, 3, 256, 256]
"""
import ctypes
import glob
import math
import os
import threading

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset

from . import _lib

SCALE = (0.08, 1.0)                      # torchvision RandomResizedCrop defaults, which datasets.py:28 uses
RATIO = (3.0 / 4.0, 4.0 / 3.0)


def sample_params(height, width, scale=SCALE, ratio=RATIO):
    """The random draws of datasets.py:27-30 for one image, from torch's global generator in torchvision's order:
    RandomResizedCrop.get_params (up to 10 attempts of area ~ U(scale) * H * W, log-uniform aspect ratio, uniform
    position; else the central crop clipped to the ratio range), then one draw per flip (p = 0.5).
    -> (top, left, crop_h, crop_w, hflip, vflip)"""
    area = height * width
    log_ratio = torch.log(torch.tensor(ratio))
    for _ in range(10):
        target_area = area * torch.empty(1).uniform_(scale[0], scale[1]).item()
        aspect_ratio = torch.exp(torch.empty(1).uniform_(log_ratio[0], log_ratio[1])).item()
        w = int(round(math.sqrt(target_area * aspect_ratio)))
        h = int(round(math.sqrt(target_area / aspect_ratio)))
        if 0 < w <= width and 0 < h <= height:
            top = torch.randint(0, height - h + 1, size=(1,)).item()
            left = torch.randint(0, width - w + 1, size=(1,)).item()
            break
    else:
        in_ratio = float(width) / float(height)
        if in_ratio < min(ratio):
            w = width
            h = int(round(w / min(ratio)))
        elif in_ratio > max(ratio):
            h = height
            w = int(round(h * max(ratio)))
        else:
            w, h = width, height
        top, left = (height - h) // 2, (width - w) // 2
    hflip = bool(torch.rand(1) < 0.5)
    vflip = bool(torch.rand(1) < 0.5)
    return top, left, h, w, hflip, vflip


def _as_u8_hwc(img):
    t = torch.from_numpy(np.ascontiguousarray(img)) if isinstance(img, np.ndarray) else img
    if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
        raise ValueError(f"images must be uint8 [H, W, 3] (decoded RGB), got {t.dtype} {tuple(t.shape)}")
    return t.contiguous()


_staging = {}                                    # (device index, thread) -> (pinned uint8 buffer, event of its last copy)


def _staging_buffer(nbytes, key):
    """grow-only pinned staging buffer; waits for the copy that last read it before handing it out again"""
    buf, ev = _staging.get(key, (None, None))
    if ev is not None:
        ev.synchronize()
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty((max(nbytes, 1 << 20) * 5 // 4,), dtype=torch.uint8, pin_memory=True)
    _staging[key] = (buf, None)
    return buf


def pack_images(images, params, crop_only=True, staging_key=None):
    """decoded images + their draws -> (pinned uint8 buffer, int64 descriptors [B, 8], max crop_h, max crop_w).
    With crop_only (default) only each image's crop rectangle is staged (the descriptor then describes a crop_h x crop_w
    image cropped in full), so the host->device copy moves the bytes the transform reads and nothing else.  Every image
    starts at a multiple of 16 bytes; descriptor layout as in include/b2c.h (b2c_resized_crop_flip)."""
    if len(images) != len(params) or not images:
        raise ValueError("need one (top, left, h, w, hflip, vflip) per image and at least one image")
    imgs = [_as_u8_hwc(i) for i in images]
    rows = []
    total = 0
    for n, (t, p) in enumerate(zip(imgs, params)):
        H, W = t.shape[0], t.shape[1]
        top, left, h, w, hf, vf = (int(v) for v in p)
        if not (0 <= top and 0 <= left and 0 < h and 0 < w and top + h <= H and left + w <= W):
            raise ValueError(f"crop box {(top, left, h, w)} outside image {n} of size {(H, W)}")
        flags = (1 if hf else 0) | (2 if vf else 0)
        if crop_only:
            imgs[n] = t[top:top + h, left:left + w]
            rows.append([total, h, w, 0, 0, h, w, flags])
            total += (h * w * 3 + 15) // 16 * 16
        else:
            rows.append([total, H, W, top, left, h, w, flags])
            total += (H * W * 3 + 15) // 16 * 16
    desc = torch.tensor(rows, dtype=torch.int64)
    if torch.cuda.is_available():
        buf = _staging_buffer(total, staging_key)[:total]
        desc = desc.pin_memory()
    else:
        buf = torch.empty((total,), dtype=torch.uint8)
    for r, t in zip(rows, imgs):
        buf[r[0]:r[0] + r[1] * r[2] * 3].view(r[1], r[2], 3).copy_(t)
    return buf, desc, int(desc[:, 5].max()), int(desc[:, 6].max())


def resized_crop_flip_packed(src, desc, size, max_crop_h, max_crop_w, check=True):
    """device uint8 buffer + device int64 [B, 8] descriptors -> CUDA fp32 [B, 3, size, size] (one launch)"""
    if not (src.is_cuda and src.dtype == torch.uint8 and desc.is_cuda and desc.dtype == torch.int64 and desc.dim() == 2
            and desc.shape[1] == 8 and desc.device == src.device):
        raise _lib.B2CError("resized_crop_flip: need a CUDA uint8 buffer and CUDA int64 [B, 8] descriptors on one device")
    _lib.require_device(src.device.index)
    B = desc.shape[0]
    out = torch.empty((B, 3, size, size), device=src.device, dtype=torch.float32)
    status = torch.zeros((1,), device=src.device, dtype=torch.int32)
    with torch.cuda.device(src.device):
        _lib.check(_lib.load().b2c_resized_crop_flip(
            ctypes.c_void_p(src.data_ptr()), src.numel(), ctypes.c_void_p(desc.contiguous().data_ptr()), B, size,
            max_crop_h, max_crop_w, ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(status.data_ptr()),
            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "resized_crop_flip")
    if check and int(status.item()):
        raise _lib.B2CError("resized_crop_flip: a descriptor lies outside its image, the buffer or the crop maxima")
    return out


def resized_crop_flip(images, params, size=256, device=None):
    """list of decoded uint8 [H, W, 3] images (numpy / CPU tensors) + their draws -> CUDA fp32 [B, 3, size, size]"""
    dev = torch.device(device if device is not None else "cuda")
    index = dev.index if dev.index is not None else torch.cuda.current_device()
    key = (index, threading.get_ident())         # one staging buffer per calling thread (nn.DataParallel-style callers)
    buf, desc, mh, mw = pack_images(images, params, staging_key=key)
    with torch.cuda.device(index):
        src, dd = buf.to(dev, non_blocking=True), desc.to(dev, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        _staging[key] = (_staging[key][0], ev)
        return resized_crop_flip_packed(src, dd, size, mh, mw)


def _decode(path):
    from PIL import Image
    return torch.from_numpy(np.asarray(Image.open(path).convert("RGB")).copy())      # datasets.py:25


class Datasets(Dataset):
    """datasets.py:14-37: every file of ``data_dir``; items are (decoded uint8 image, draws) for `collate_device`"""

    def __init__(self, data_dir, image_size=256):
        self.data_dir = data_dir
        self.image_size = image_size
        if not os.path.exists(data_dir):
            raise Exception(f"[!] {self.data_dir} not exitd")
        self.image_path = sorted(glob.glob(os.path.join(self.data_dir, "*.*")))

    def __getitem__(self, item):
        img = _decode(self.image_path[item])
        return img, sample_params(img.shape[0], img.shape[1])

    def __len__(self):
        return len(self.image_path)


class TestKodakDataset(Dataset):
    """datasets.py:63-80: ``ToTensor()`` only, so items are fp32 ``[3, H, W]`` tensors in [0, 1] exactly as the
    reference returns them (a default DataLoader over this dataset feeds the model the reference's dtype and layout;
    ``x / 255`` below is ToTensor's own arithmetic, bit for bit).  ``raw=True`` returns uint8 ``[3, H, W]`` instead,
    for device-side conversion: ``inference.HostPipeline.compress_u8`` moves a quarter of the bytes over PCIe and
    divides on the GPU with identical results."""

    def __init__(self, data_dir, raw=False):
        self.data_dir = data_dir
        self.raw = raw
        if not os.path.exists(data_dir):
            raise Exception(f"[!] {self.data_dir} not exitd")
        self.image_path = sorted(glob.glob(os.path.join(self.data_dir, "*.*")))

    def __getitem__(self, item):
        chw = _decode(self.image_path[item]).permute(2, 0, 1).contiguous()
        return chw if self.raw else chw.to(torch.float32).div(255)

    def __len__(self):
        return len(self.image_path)


class DeviceLoader:
    """Iterates a torch DataLoader of (image, draws) items and yields device batches from one kernel launch each."""

    def __init__(self, dataset, batch_size, shuffle, image_size, device=None, num_workers=0):
        self.image_size = image_size
        self.device = device
        self.loader = DataLoader(dataset, batch_size=batch_size, shuffle=shuffle, num_workers=num_workers,
                                 collate_fn=lambda items: items)
        self.dataset = dataset

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        for items in self.loader:
            yield resized_crop_flip([i[0] for i in items], [i[1] for i in items], self.image_size, self.device)


def get_loader(train_data_dir, test_data_dir, image_size, batch_size):
    """datasets.py:40-51"""
    return (DeviceLoader(Datasets(train_data_dir, image_size), batch_size, True, image_size),
            DeviceLoader(Datasets(test_data_dir, image_size), batch_size, False, image_size))


def get_train_loader(train_data_dir, image_size, batch_size):
    """datasets.py:54-61 (including its fixed shuffle seed)"""
    train_dataset = Datasets(train_data_dir, image_size)
    torch.manual_seed(3334)
    return train_dataset, DeviceLoader(train_dataset, batch_size, True, image_size)
