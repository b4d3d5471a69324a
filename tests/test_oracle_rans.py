"""CPU checks of the bitstream oracle (oracle/rans_oracle.c, parity unpinned: the reference has no bitstream) and of the
host-side container logic of iclr_17_compression_b200.bitstream (no CUDA calls)."""
import numpy as np
import pytest

from oracle import rans_oracle as R


def _tables(rng, C, L, prec, peak=4.0):
    A = 2 * L + 2
    p = rng.random((C, A)) + 0.05
    p[:, L] += peak
    p /= p.sum(1, keepdims=True)
    f = 1 + np.floor(p * (2 ** prec - A)).astype(np.int64)
    f[np.arange(C), f.argmax(1)] += 2 ** prec - f.sum(1)
    return f


def test_single_symbol_known_answer():
    """one latent, worked by hand: L = 1 (A = 4), prec = 10, f = (256, 512, 128, 128), z = 0 -> s = 1, start = 256.
    x0 = 65536 < 2^22 * 512: no word; x = (65536 // 512 << 10) + 65536 % 512 + 256 = 131328 = 0x00020100.
    stream = lane 0's state (00 01 02 00 little endian), then 31 untouched states 0x00010000."""
    f = np.array([[256, 512, 128, 128]])
    data, off, bb, st = R.encode(np.zeros((1, 1, 1, 1), np.float32), R.pack_freq(f), 1, 10, 32)
    assert st == 0 and bb.tolist() == [128] and off.tolist() == [0, 128]
    assert data[:4].tolist() == [0x00, 0x01, 0x02, 0x00]
    assert data[4:].reshape(31, 4).tolist() == [[0, 0, 1, 0]] * 31
    z, st = R.decode(data, off, R.pack_freq(f), 1, 10, 1, 1, 1, 1, 32)
    assert st == 0 and z.item() == 0.0


def test_escape_known_answer():
    """z = 5 with L = 1: raw word first (reverse order): emit 0x0000, x = 1 << 16 | (5 + 32768) = 0x00018005; then the
    escape symbol s = 3 (f = 128, start = 896): x < 2^22 * 128 = 2^29: no word;
    x = (0x18005 // 128 << 10) + 0x18005 % 128 + 896 = (768 << 10) + 5 + 896 = 787333 = 0x000C0385; one word 0x0000 follows."""
    f = np.array([[256, 512, 128, 128]])
    data, off, bb, st = R.encode(np.full((1, 1, 1, 1), 5, np.float32), R.pack_freq(f), 1, 10, 32)
    assert st == 0 and bb.tolist() == [130]
    assert data[:4].tolist() == [0x85, 0x03, 0x0C, 0x00] and data[128:].tolist() == [0, 0]


@pytest.mark.parametrize("shape,L,prec,bs", [((3, 8, 5, 7), 3, 12, 32), ((3, 8, 5, 7), 3, 12, 96), ((2, 16, 16, 16), 31, 14, 4096),
                                             ((1, 4, 3, 3), 1, 10, 32), ((2, 128, 4, 4), 31, 15, 512)])
def test_round_trip_and_size(shape, L, prec, bs):
    rng = np.random.default_rng(sum(shape) + L)
    f = _tables(rng, shape[1], L, prec)
    z = np.rint(rng.normal(0, 2.0, shape)).astype(np.float32)
    z.reshape(-1)[0] = -300                                   # escapes
    z.reshape(-1)[-1] = 32767
    data, off, bb, st = R.encode(z, R.pack_freq(f), L, prec, bs)
    assert st == 0 and off[-1] == len(data) == bb.sum()
    z2, st2 = R.decode(data, off, R.pack_freq(f), L, prec, *shape, bs)
    assert st2 == 0 and np.array_equal(z, z2)
    # size: the 32 states cost 128 bytes per block but carry 16 bits each of payload; the rest is within 1 % of ideal
    ideal = R.ideal_bits(z, f, L, prec) / 8
    assert ideal <= len(data) <= ideal * 1.01 + 128 * len(bb) + 4


def test_out_of_range_and_corruption_are_flagged():
    rng = np.random.default_rng(5)
    f = _tables(rng, 4, 3, 12)
    z = np.zeros((1, 4, 8, 8), np.float32)
    z[0, 0, 0, 0] = 40000
    _, _, _, st = R.encode(z, R.pack_freq(f), 3, 12, 64)
    assert st & 1
    z = np.rint(rng.normal(0, 2, (1, 4, 8, 8))).astype(np.float32)
    data, off, bb, st = R.encode(z, R.pack_freq(f), 3, 12, 64)
    assert st == 0
    bad = data.copy()
    bad[3] ^= 0x40                                            # a state bit
    _, st2 = R.decode(bad, off, R.pack_freq(f), 3, 12, 1, 4, 8, 8, 64)
    assert st2 & 4


def test_frequencies64_are_valid_tables():
    import torch
    g = torch.Generator().manual_seed(3003)
    params = [torch.randn(16, generator=g).numpy() * s for s in (1, .5, .5, 1, .5, .5, 1, .5, .5, 1, .5)]
    for L, prec in ((31, 14), (7, 10), (126, 15)):
        f = R.frequencies64(params, L, prec)
        assert f.shape == (16, 2 * L + 2) and (f >= 1).all() and (f.sum(1) == 2 ** prec).all()
    # with a wide table the probabilities must match the reference formula p = CDF(z+.5) - CDF(z-.5) (model.py:71-72)
    f = R.frequencies64(params, 31, 14).astype(np.float64) / 2 ** 14
    zs = np.arange(-31, 32, dtype=np.float64)[None].repeat(16, 0)
    p = R.cdf64(params, zs + .5) - R.cdf64(params, zs - .5)
    assert np.abs(f[:, :-1] - p).max() < 0.01


def test_container_parse_rejects_malformed():
    from iclr_17_compression_b200 import bitstream as BS
    head = BS._HEAD.pack(BS.MAGIC, BS.VERSION, 14, 31, 0, 128, 2, 2, 32768, 1)
    good = head + np.array([130], "<u4").tobytes() + bytes(130)
    hd, bb, payload = BS.parse_stream(good)
    assert hd["C"] == 128 and bb.tolist() == [130] and len(payload) == 130
    for bad in (good[:10], b"XXXX" + good[4:], good[:-1], head):
        with pytest.raises(ValueError):
            BS.parse_stream(bad)


def test_codec_file_container_and_padding():
    """host side of codec.py: file header round trip, consistency with the embedded stream, edge padding"""
    import struct
    from iclr_17_compression_b200 import bitstream as BS
    from iclr_17_compression_b200 import codec
    nblk = 2
    stream = BS._HEAD.pack(BS.MAGIC, BS.VERSION, 14, 31, 0, 128, 4, 5, 8192, nblk) + struct.pack("<II", 3, 1) + b"abcd"
    data = codec.pack_file(stream, 128, 50, 70)
    assert codec.unpack_file(data) == (128, 50, 70, stream)
    for bad in (data[:10], b"XXXX" + data[4:], codec.pack_file(stream, 64, 50, 70), codec.pack_file(stream, 128, 65, 70),
                codec.pack_file(stream, 128, 48, 70), codec.pack_file(stream[:-1], 128, 50, 70)):
        with pytest.raises(ValueError):
            codec.unpack_file(bad)
    img = np.arange(5 * 7 * 3, dtype=np.uint8).reshape(5, 7, 3)
    p = codec.pad_to_16(img)
    assert p.shape == (16, 16, 3) and np.array_equal(p[:5, :7], img)
    assert np.array_equal(p[5:, :7], np.broadcast_to(img[4:5], (11, 7, 3))) and np.array_equal(p[:, 7:], np.broadcast_to(p[:, 6:7], (16, 9, 3)))
    assert codec.pad_to_16(np.zeros((32, 48, 3), np.uint8)).shape == (32, 48, 3)
