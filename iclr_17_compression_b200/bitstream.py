"""Real bitstream for the integer latents.

The reference only *estimates* the size of the latents (``model.py:71-74``) and approximates files with
``gzip(np.packbits(latent == 0))`` (``NewTests/testModelRecon.py:80``).  Here the same per-channel probabilities
``CDF(z + .5) - CDF(z - .5)`` of the ``BitEstimator`` drive a warp-interleaved range-ANS coder on the GPU (libb2c
``b2c_rans_*``, format in ``csrc/b2c_rans.cu``): ``compress`` turns images into one byte string per image,
``decompress`` turns them back into reconstructions.  CUDA only; there is no CPU path.

    tables  = RansTables.from_bit_estimator(model.bitEstimator)
    streams = compress(model, x, tables)            # list[bytes], one per image
    recon   = decompress(model, streams, tables)    # == model(x)[0]

Per-image container (little endian): ``b"B2CR"``, u8 version, u8 prec, u8 half_range, u8 0, u32 C, u32 h, u32 w (latent
dims), u32 block_symbols, u32 nblk, u32 block_bytes[nblk], then the nblk block streams.
"""
import struct

import numpy as np
import torch

from . import _lib
from . import functional as B2F

MAGIC = b"B2CR"
VERSION = 1
_HEAD = struct.Struct("<4sBBBBIIIII")

ST_RANGE, ST_CAPACITY, ST_CORRUPT = 1, 2, 4


class RansTables:
    """Per-channel frequency tables of the coder, on the device: ``freq_cum`` int32 [C, A] holding
    ``f | start << 16`` (A = 2 * half_range + 2, the last symbol is the escape) and ``slot_sym`` uint8 [C, 2**prec]."""

    def __init__(self, freq_cum, slot_sym, half_range, prec):
        self.freq_cum, self.slot_sym, self.half_range, self.prec = freq_cum, slot_sym, int(half_range), int(prec)
        self.C = freq_cum.shape[0]

    @classmethod
    def from_bit_estimator(cls, bit_estimator, half_range=31, prec=14):
        """tables from the BitEstimator's parameters (bitEstimator.py:13-18,38-42), built by one kernel launch"""
        params = [B2F._prep(p.detach().reshape(-1), "bit-estimator parameter") for p in bit_estimator.param_list()]
        C, dev = params[0].numel(), params[0].device
        A, M = 2 * half_range + 2, 1 << prec
        fc = torch.empty((C, A), device=dev, dtype=torch.int32)
        ss = torch.empty((C, M), device=dev, dtype=torch.uint8)
        pp, keep = _lib.ptr_array(params)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().b2c_rans_build_tables(pp, C, half_range, prec, B2F._p(fc), B2F._p(ss), B2F._stream()),
                       "rans_build_tables")
        del keep
        return cls(fc, ss, half_range, prec)

    @classmethod
    def from_frequencies(cls, freq, prec, device):
        """tables from explicit integer frequencies [C, A] (every entry >= 1, rows summing to 2**prec)"""
        freq = np.asarray(freq, dtype=np.int64)
        C, A = freq.shape
        M = 1 << prec
        if A % 2 or A < 4 or (freq < 1).any() or (freq.sum(1) != M).any():
            raise ValueError("frequencies must be [C, 2L+2], all >= 1, each row summing to 2**prec")
        start = np.concatenate([np.zeros((C, 1), np.int64), np.cumsum(freq, 1)[:, :-1]], 1)
        fc = (freq | (start << 16)).astype(np.uint32).view(np.int32)
        ss = np.repeat(np.tile(np.arange(A, dtype=np.uint8), C), freq.reshape(-1)).reshape(C, M)
        dev = torch.device(device)
        _lib.require_device(dev.index if dev.index is not None else torch.cuda.current_device())
        return cls(torch.from_numpy(fc).to(dev), torch.from_numpy(ss).to(dev), (A - 2) // 2, prec)

    def frequencies(self):
        """[C, A] numpy int64 (device -> host copy)"""
        return (self.freq_cum.cpu().numpy().view(np.uint32) & 0xFFFF).astype(np.int64)


def _status(status, what):
    s = int(status.item())
    if s & ST_RANGE:
        raise _lib.B2CError(f"{what}: latent value outside [-32767, 32767] (or not finite)")
    if s & ST_CAPACITY:
        raise _lib.B2CError(f"{what}: output buffer too small")
    if s & ST_CORRUPT:
        raise _lib.B2CError(f"{what}: corrupt or truncated stream (final state / length check failed)")


def encode_latents_raw(z, tables, block_symbols=32768, out_capacity=None):
    """z: fp32 CUDA [B, C, h, w] holding integers -> (out uint8 [capacity], offsets int64 [B*nblk+1],
    block_bytes int32 [B*nblk], status int32 [1]) on the device, nothing synchronised."""
    z = B2F._prep(z, "latent")
    B, C, h, w = z.shape
    if C != tables.C:
        raise _lib.B2CError(f"latent has {C} channels, tables {tables.C}")
    HW = h * w
    lib = _lib.load()
    nblk = lib.b2c_rans_num_blocks(C, HW, block_symbols)
    sbytes = lib.b2c_rans_scratch_bytes(B, C, HW, block_symbols)
    if nblk <= 0 or sbytes == 0:
        raise _lib.B2CError("rans_encode: bad geometry (block_symbols must be a multiple of 32, >= 32)")
    dev = z.device
    scratch = torch.empty((sbytes,), device=dev, dtype=torch.uint8)
    cap = int(out_capacity) if out_capacity is not None else sbytes      # worst case unless the caller knows better
    out = torch.empty((cap + (cap & 1),), device=dev, dtype=torch.uint8)
    block_bytes = torch.empty((B * nblk,), device=dev, dtype=torch.int32)
    offsets = torch.empty((B * nblk + 1,), device=dev, dtype=torch.int64)
    status = torch.zeros((1,), device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        _lib.check(lib.b2c_rans_encode(B2F._p(z), B2F._p(tables.freq_cum), tables.half_range, tables.prec, B, C, HW,
                                       block_symbols, B2F._p(scratch), sbytes, B2F._p(block_bytes), B2F._p(offsets),
                                       B2F._p(out), cap, B2F._p(status), B2F._stream()), "rans_encode")
    return out, offsets, block_bytes, status


def decode_latents_raw(data, offsets, tables, B, C, h, w, block_symbols=32768):
    """inverse of encode_latents_raw: device uint8 bytes + int64 offsets -> (z fp32 [B, C, h, w], status)"""
    if not (data.is_cuda and data.dtype == torch.uint8 and offsets.is_cuda and offsets.dtype == torch.int64):
        raise _lib.B2CError("rans_decode: data must be CUDA uint8 and offsets CUDA int64")
    lib = _lib.load()
    nblk = lib.b2c_rans_num_blocks(C, h * w, block_symbols)
    if offsets.numel() != B * nblk + 1:
        raise _lib.B2CError(f"rans_decode: expected {B * nblk + 1} offsets, got {offsets.numel()}")
    _lib.require_device(data.device.index)
    z = torch.empty((B, C, h, w), device=data.device, dtype=torch.float32)
    status = torch.zeros((1,), device=data.device, dtype=torch.int32)
    with torch.cuda.device(data.device):
        _lib.check(lib.b2c_rans_decode(B2F._p(data.contiguous()), B2F._p(offsets.contiguous()), B2F._p(tables.freq_cum),
                                       B2F._p(tables.slot_sym), tables.half_range, tables.prec, B, C, h * w, block_symbols,
                                       B2F._p(z), B2F._p(status), B2F._stream()), "rans_decode")
    return z, status


def encode_latents(z, tables, block_symbols=32768):
    """z [B, C, h, w] (integers in fp32, CUDA) -> list of B byte strings (container in the module docstring)"""
    B, C, h, w = z.shape
    out, offsets, block_bytes, status = encode_latents_raw(z, tables, block_symbols)
    nb = block_bytes.numel()
    meta = torch.cat([offsets, block_bytes.to(torch.int64), status.to(torch.int64)]).cpu().numpy()   # one small D2H
    if int(meta[-1]):
        _status(status, "rans_encode")
    off, bb = meta[:nb + 1], meta[nb + 1:2 * nb + 1].astype("<u4")
    payload = out[:int(off[-1])].cpu().numpy().tobytes()
    nblk = bb.size // B
    streams = []
    for n in range(B):
        head = _HEAD.pack(MAGIC, VERSION, tables.prec, tables.half_range, 0, C, h, w, block_symbols, nblk)
        streams.append(head + bb[n * nblk:(n + 1) * nblk].tobytes() + payload[int(off[n * nblk]):int(off[(n + 1) * nblk])])
    return streams


def parse_stream(stream):
    """-> (dict header, block_bytes uint32 [nblk], payload bytes); raises ValueError on a malformed container"""
    if len(stream) < _HEAD.size:
        raise ValueError("stream shorter than its header")
    magic, ver, prec, L, _, C, h, w, bs, nblk = _HEAD.unpack_from(stream)
    if magic != MAGIC or ver != VERSION:
        raise ValueError("not a B2CR version-1 stream")
    if len(stream) < _HEAD.size + 4 * nblk:
        raise ValueError("stream shorter than its block table")
    bb = np.frombuffer(stream, dtype="<u4", count=nblk, offset=_HEAD.size)
    payload = stream[_HEAD.size + 4 * nblk:]
    if int(bb.sum()) != len(payload):
        raise ValueError(f"payload is {len(payload)} bytes, block table says {int(bb.sum())}")
    return dict(prec=prec, half_range=L, C=C, h=h, w=w, block_symbols=bs, nblk=nblk), bb, payload


def decode_latents(streams, tables, device=None):
    """list of byte strings with identical geometry -> z fp32 [B, C, h, w] on the tables' device"""
    if not streams:
        raise ValueError("no streams")
    parsed = [parse_stream(s) for s in streams]
    hd = parsed[0][0]
    for p in parsed[1:]:
        if p[0] != hd:
            raise ValueError("streams of one batch must share geometry and coder settings")
    if (hd["prec"], hd["half_range"], hd["C"]) != (tables.prec, tables.half_range, tables.C):
        raise ValueError(f"stream was coded with prec={hd['prec']} half_range={hd['half_range']} C={hd['C']}, tables have "
                         f"prec={tables.prec} half_range={tables.half_range} C={tables.C}")
    dev = torch.device(device) if device is not None else tables.freq_cum.device
    bb = np.concatenate([p[1] for p in parsed]).astype(np.int64)
    off = np.concatenate([[0], np.cumsum(bb)])
    payload = b"".join(p[2] for p in parsed)
    data = torch.frombuffer(bytearray(payload + b"\0\0"), dtype=torch.uint8).to(dev)
    z, status = decode_latents_raw(data, torch.from_numpy(off).to(dev), tables, len(streams), hd["C"], hd["h"], hd["w"],
                                   hd["block_symbols"])
    _status(status, "rans_decode")
    return z


@torch.no_grad()
def compress(model, x, tables=None, block_symbols=32768):
    """images [B, 3, H, W] in [0, 1] -> (streams, image_bits): analysis transform, round (model.py:56), range coder.
    image_bits are the reference's ESTIMATE for the same latents (model.py:71-74), for comparison with 8 * len(stream)."""
    if model.training:
        raise RuntimeError("compress() needs model.eval(): the training quantizer adds noise (model.py:53-54)")
    x = B2F._prep(x, "input image")
    tables = tables if tables is not None else RansTables.from_bit_estimator(model.bitEstimator)
    params = [p.detach().reshape(-1) for p in model.bitEstimator.param_list()]
    if model._use_tc(x):
        z, _, image_bits = B2F.quant_entropy_infer(model.encode_tc(x), params, lut=model._packed_weights()["bits_lut"])
    else:
        z, _, image_bits, _ = model.bitEstimator.bits(model._enc(x), None)
    return encode_latents(z, tables, block_symbols), image_bits


@torch.no_grad()
def decompress(model, streams, tables=None):
    """byte strings -> clamped reconstruction [B, 3, 16h, 16w] (synthesis transform of the decoded latents, model.py:57-59)"""
    tables = tables if tables is not None else RansTables.from_bit_estimator(model.bitEstimator)
    z = decode_latents(streams, tables)
    if model._use_tc(z):
        return model.decode_tc(z).clamp_(0.0, 1.0)
    return model._dec(z).clamp(0.0, 1.0)
