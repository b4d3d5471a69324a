"""TEST INFRASTRUCTURE - Python side of the bitstream oracle (oracle/rans_oracle.c) plus the fp64 restatement of the
frequency-table rule of csrc/b2c_rans.cu.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
import this.  PARITY UNPINNED: the reference has no bitstream (see rans_oracle.c); the probabilities are the
reference's (models/bitEstimator.py:20-25,38-42, model.py:71-72)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "rans_oracle.c")
LIB = os.path.join(_HERE, "_build", "librans_oracle.so")


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.run(["gcc", "-O2", "-Wall", "-shared", "-fPIC", SRC, "-o", LIB, "-lm"], check=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        vp, i = ctypes.c_void_p, ctypes.c_int
        _lib.rans_oracle_encode.argtypes = [vp, vp, i, i, i, i, i, i, vp, vp, vp, ctypes.c_size_t, vp]
        _lib.rans_oracle_decode.argtypes = [vp, vp, vp, i, i, i, i, i, i, vp, vp]
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def pack_freq(freq):
    """[C, A] integer frequencies -> uint32 f | start << 16"""
    freq = np.asarray(freq, dtype=np.int64)
    start = np.concatenate([np.zeros((freq.shape[0], 1), np.int64), np.cumsum(freq, 1)[:, :-1]], 1)
    return np.ascontiguousarray((freq | (start << 16)).astype(np.uint32))


def encode(z, freq_cum, half_range, prec, block_symbols):
    """z float32 [B, C, h, w] -> (bytes uint8, offsets uint64 [B*nblk+1], block_bytes uint32, status)"""
    z = np.ascontiguousarray(z, dtype=np.float32)
    fc = np.ascontiguousarray(freq_cum, dtype=np.uint32)
    B, C, h, w = z.shape
    T = C * h * w
    nblk = (T + block_symbols - 1) // block_symbols
    cap = B * nblk * (2 * block_symbols + 64) * 2
    out = np.zeros(cap, np.uint8)
    offsets = np.zeros(B * nblk + 1, np.uint64)
    bb = np.zeros(B * nblk, np.uint32)
    status = np.zeros(1, np.int32)
    rc = lib().rans_oracle_encode(_ptr(z), _ptr(fc), half_range, prec, B, C, h * w, block_symbols, _ptr(bb), _ptr(offsets),
                                  _ptr(out), cap, _ptr(status))
    assert rc == 0
    return out[:int(offsets[-1])].copy(), offsets, bb, int(status[0])


def decode(data, offsets, freq_cum, half_range, prec, B, C, h, w, block_symbols):
    data = np.ascontiguousarray(data, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    fc = np.ascontiguousarray(freq_cum, dtype=np.uint32)
    z = np.zeros((B, C, h, w), np.float32)
    status = np.zeros(1, np.int32)
    rc = lib().rans_oracle_decode(_ptr(data), _ptr(offsets), _ptr(fc), half_range, prec, B, C, h * w, block_symbols, _ptr(z),
                                  _ptr(status))
    assert rc == 0
    return z, int(status[0])


def cdf64(params, v):
    """BitEstimator CDF in float64: params = 11 arrays [C] in libb2c order (f1.h, f1.b, f1.a, ..., f4.h, f4.b);
    v [C, K] -> [C, K]  (bitEstimator.py:20-25: t = x * softplus(h) + b ; x' = t + tanh(t) * tanh(a) ; last: sigmoid)"""
    p = [np.asarray(q, dtype=np.float64).reshape(-1, 1) for q in params]
    sp = lambda h: np.where(h > 20, h, np.log1p(np.exp(np.minimum(h, 20))))
    for k in range(3):
        t = v * sp(p[3 * k]) + p[3 * k + 1]
        v = t + np.tanh(t) * np.tanh(p[3 * k + 2])
    return 1.0 / (1.0 + np.exp(-(v * sp(p[9]) + p[10])))


def frequencies64(params, half_range, prec):
    """the table rule of csrc/b2c_rans.cu in float64: f = 1 + floor(p * (2^prec - A)), escape = both tails, remainder to
    the (first) most probable symbol -> int64 [C, A]"""
    L, A, M = half_range, 2 * half_range + 2, 1 << prec
    C = np.asarray(params[0]).size
    zs = np.arange(-L, L + 1, dtype=np.float64)[None, :].repeat(C, 0)
    p = cdf64(params, zs + 0.5) - cdf64(params, zs - 0.5)
    tails = cdf64(params, np.full((C, 1), -L - 0.5)) + (1.0 - cdf64(params, np.full((C, 1), L + 0.5)))
    p = np.clip(np.concatenate([p, tails], 1), 0.0, 1.0)
    f = 1 + np.floor(p * (M - A)).astype(np.int64)
    best = f.argmax(1)
    f[np.arange(C), best] += M - f.sum(1)
    return f


def ideal_bits(z, freq, half_range, prec):
    """sum over elements of -log2(f / 2^prec) (+16 per escape): what a perfect coder spends with these tables"""
    z = np.asarray(z)
    B, C = z.shape[:2]
    zi = np.rint(z).astype(np.int64).reshape(B, C, -1)
    esc = np.abs(zi) > half_range
    s = np.where(esc, 2 * half_range + 1, zi + half_range)
    f = np.take_along_axis(np.broadcast_to(freq[None], (B,) + freq.shape), s, 2)
    return float((prec - np.log2(f)).sum() + 16.0 * esc.sum())
