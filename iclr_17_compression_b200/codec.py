"""Image file <-> bitstream file with the ImageCompressor codec ("real bitstream + on-disk latent
format"; the reference only estimates the size, ``model.py:71-74``, and writes no files).

    python -m iclr_17_compression_b200.codec encode photo.png photo.b2c [--checkpoint iter_x.pth.tar] [-N 128]
    python -m iclr_17_compression_b200.codec decode photo.b2c photo_rec.png [--checkpoint ...] [-N 128]

A ``.b2c`` file is one self-describing stream of ``bitstream`` (container in that module's docstring) preceded by a
16-byte file header: ``b"B2CF"``, u16 version, u16 out_channel_N, u32 image height, u32 image width.  The codec
needs H and W to be multiples of 16 (the reference's own constraint, ``model.py:48,54``); ``encode`` pads on the
bottom / right by edge replication and ``decode`` crops back to the stored size.  Both directions must use the same
weights.  CUDA only (libb2c); the image decode / encode of the container formats is PIL on the host.
"""
import argparse
import struct

import numpy as np
import torch

from . import bitstream as BS
from .model import ImageCompressor, load_model

FILE_MAGIC = b"B2CF"
FILE_VERSION = 1
_FHEAD = struct.Struct("<4sHHII")


def pack_file(stream, out_channel_N, height, width):
    return _FHEAD.pack(FILE_MAGIC, FILE_VERSION, out_channel_N, height, width) + stream


def unpack_file(data):
    """-> (out_channel_N, height, width, stream); ValueError on a malformed file"""
    if len(data) < _FHEAD.size:
        raise ValueError("file shorter than its header")
    magic, ver, n, h, w = _FHEAD.unpack_from(data)
    if magic != FILE_MAGIC or ver != FILE_VERSION:
        raise ValueError("not a B2CF version-1 file")
    stream = data[_FHEAD.size:]
    hd, _, _ = BS.parse_stream(stream)
    if hd["C"] != n or hd["h"] * 16 < h or hd["w"] * 16 < w or (hd["h"] - 1) * 16 >= h or (hd["w"] - 1) * 16 >= w:
        raise ValueError(f"file header ({n} channels, {h} x {w}) does not match its stream "
                         f"({hd['C']} channels, latent {hd['h']} x {hd['w']})")
    return n, h, w, stream


def pad_to_16(img_u8):
    """uint8 [H, W, 3] -> uint8 [H16, W16, 3], edge-replicated on the bottom / right"""
    h, w = img_u8.shape[:2]
    return np.pad(img_u8, ((0, -h % 16), (0, -w % 16), (0, 0)), mode="edge")


def build_model(out_channel_N=128, checkpoint=None, device="cuda"):
    torch.manual_seed(234)                                     # train.py:39 default seed when no checkpoint is given
    model = ImageCompressor(out_channel_N)
    if checkpoint:
        load_model(model, checkpoint)
    return model.to(device).eval()


def encode_image(model, img_u8, tables=None):
    """uint8 [H, W, 3] -> bytes of a .b2c file"""
    h, w = img_u8.shape[:2]
    x = torch.from_numpy(pad_to_16(img_u8)).to(next(model.parameters()).device)
    x = x.permute(2, 0, 1).unsqueeze(0).to(torch.float32).div(255)          # ToTensor (datasets.py:74)
    streams, _ = BS.compress(model, x.contiguous(), tables)
    return pack_file(streams[0], model.out_channel_N, h, w)


def decode_image(model, data, tables=None):
    """bytes of a .b2c file -> uint8 [H, W, 3]"""
    n, h, w, stream = unpack_file(data)
    if n != model.out_channel_N:
        raise ValueError(f"file was coded with out_channel_N={n}, the model has {model.out_channel_N}")
    rec = BS.decompress(model, [stream], tables)[0, :, :h, :w]
    return rec.mul(255).round().clamp(0, 255).to(torch.uint8).permute(1, 2, 0).contiguous().cpu().numpy()


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m iclr_17_compression_b200.codec", description=__doc__.split("\n")[0])
    ap.add_argument("mode", choices=["encode", "decode"])
    ap.add_argument("src")
    ap.add_argument("dst")
    ap.add_argument("--checkpoint", default=None, help="file written by save_model (model.py:18-20)")
    ap.add_argument("-N", "--out-channel-n", type=int, default=128)
    args = ap.parse_args(argv)
    from PIL import Image
    model = build_model(args.out_channel_n, args.checkpoint)
    if args.mode == "encode":
        img = np.asarray(Image.open(args.src).convert("RGB"))
        data = encode_image(model, img)
        with open(args.dst, "wb") as f:
            f.write(data)
        print(f"{args.src}: {img.shape[1]} x {img.shape[0]} -> {len(data)} bytes "
              f"({8 * len(data) / (img.shape[0] * img.shape[1]):.4f} bpp)")
    else:
        with open(args.src, "rb") as f:
            img = decode_image(model, f.read())
        Image.fromarray(img, "RGB").save(args.dst)
        print(f"{args.src} -> {args.dst}: {img.shape[1]} x {img.shape[0]}")


if __name__ == "__main__":
    main()
