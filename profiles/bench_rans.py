"""Bitstream timings on one B200: table build, encode, decode (CUDA events, L2 flushed between
iterations), stream size vs the reference's estimate, and the CPU oracle (oracle/rans_oracle.c, 1 core) on the same
latents.  python profiles/bench_rans.py > gpurun_out/rans.json"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch

import iclr_17_compression_b200 as P
from iclr_17_compression_b200 import bitstream as BS
from golden_util import weights as W
from oracle import balle17_oracle as O
from oracle import rans_oracle as R


def timed(fn, iters=20, warm=3):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for i in range(warm + iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


res = {}
m = P.ImageCompressor(128)
m.load_state_dict(W(128, "W1"))
m = m.cuda().eval()
t_tab = timed(lambda: BS.RansTables.from_bit_estimator(m.bitEstimator))
tables = BS.RansTables.from_bit_estimator(m.bitEstimator)
res["table_build_ms"] = t_tab
for name, (B, H, Wd), kind in (("cfg2_24x512x768_S", (24, 512, 768), "S"), ("cfg2_24x512x768_U", (24, 512, 768), "U"),
                              ("cfg5_8x1536x2048_S", (8, 1536, 2048), "S")):
    x = O.synth_images(kind, B, H, Wd, 1003).cuda()
    with torch.no_grad():
        _, z, bpp = m(x)
    for bs in (8192, 32768, 131072):
        out, off, bb, st = BS.encode_latents_raw(z, tables, bs)
        torch.cuda.synchronize()
        assert st.item() == 0
        total = int(off[-1])
        t_enc = timed(lambda: BS.encode_latents_raw(z, tables, bs, out_capacity=total))
        t_dec = timed(lambda: BS.decode_latents_raw(out, off, tables, *z.shape, bs))
        z2, st2 = BS.decode_latents_raw(out, off, tables, *z.shape, bs)
        assert st2.item() == 0 and torch.equal(z2, z)
        est_bits = bpp.item() * B * H * Wd
        r = {"latents": z.numel(), "block_symbols": bs, "blocks": int(bb.numel()), "encode_ms": t_enc, "decode_ms": t_dec,
             "encode_Gsym_s": z.numel() / t_enc / 1e6, "decode_Gsym_s": z.numel() / t_dec / 1e6,
             "stream_bytes": total, "estimate_bytes": est_bits / 8, "stream_over_estimate": total * 8 / est_bits,
             "bpp_estimate": bpp.item(), "bpp_stream": total * 8 / (B * H * Wd),
             "escapes": int((z.abs() > tables.half_range).sum())}
        if bs == 32768 and name.startswith("cfg2"):
            zc = z[:4].cpu().numpy()
            fc = tables.freq_cum.cpu().numpy().view(np.uint32)
            t0 = time.perf_counter()
            data, offs, _, _ = R.encode(zc, fc, tables.half_range, tables.prec, bs)
            t1 = time.perf_counter()
            R.decode(data, offs, fc, tables.half_range, tables.prec, *zc.shape, bs)
            t2 = time.perf_counter()
            r["cpu_oracle_1core"] = {"latents": zc.size, "encode_Msym_s": zc.size / (t1 - t0) / 1e6,
                                      "decode_Msym_s": zc.size / (t2 - t1) / 1e6}
        if bs == 8192:
            # the public calls, host clock, everything included: network + coder + device->host copy of the bytes, and back
            def wall(fn, n=5):
                fn()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(n):
                    o = fn()
                torch.cuda.synchronize()
                return (time.perf_counter() - t0) / n * 1e3, o
            t_c, (streams, _) = wall(lambda: BS.compress(m, x, tables, bs))
            t_d, rec = wall(lambda: BS.decompress(m, streams, tables))
            with torch.no_grad():
                assert torch.equal(rec, m(x)[0])
            r["public_compress_ms"] = t_c
            r["public_decompress_ms"] = t_d
            r["public_compress_Mpix_s"] = B * H * Wd / t_c / 1e3
            r["public_decompress_Mpix_s"] = B * H * Wd / t_d / 1e3
            r["file_bytes"] = sum(len(s_) for s_ in streams)
        res[f"{name}_bs{bs}"] = r
print(json.dumps(res, indent=1))
