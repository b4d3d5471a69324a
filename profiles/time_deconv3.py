"""A/B of the last synthesis layer: DECONV3 (3x3-shift form, N = 48) vs DECONV3X (x-folded, N = 112) at the bench
shape (24 x 128 x 192 x 128 bf16 in, 24 x 3 x 512 x 768 fp32 out) and the training shape (8 x 64 x 64), CUDA events,
L2 flushed between launches; optional clock64 timeline of CTA 0."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iclr_17_compression_b200 import _lib, tc

lib = _lib.load()
lib.b2c_tc_debug_set_trace.argtypes = [ctypes.c_void_p]
g = torch.Generator().manual_seed(7)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def timeit_b2b(fn, n=20):
    """back-to-back launches (host latency hidden behind the queue; in + out exceed the L2 at the bench shape)"""
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush.zero_()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def trace(fn):
    buf = torch.zeros(3 * 64, dtype=torch.int64, device="cuda")
    lib.b2c_tc_debug_set_trace(ctypes.c_void_p(buf.data_ptr()))
    fn()
    torch.cuda.synchronize()
    lib.b2c_tc_debug_set_trace(ctypes.c_void_p(0))
    t = buf.cpu().view(3, 64)
    t0 = min(int(v) for v in t.flatten() if v > 0)
    for role, nm in enumerate(["producer", "mma", "worker"]):
        print(f"    {nm:8s}", [int(v) - t0 for v in t[role] if v > 0][:40])


for (B, H, W) in [(24, 128, 192), (8, 64, 64)]:
    x = torch.randn(B, H, W, 128, generator=g).cuda().to(torch.bfloat16)
    w = (torch.randn(128, 3, 9, 9, generator=g) * 0.05).cuda()
    b = torch.randn(3, generator=g).cuda()
    img = torch.rand(B, 3, 4 * H, 4 * W, generator=g).cuda()
    for kind, nm in [(tc.TC_DECONV3, "DECONV3 "), (tc.TC_DECONV3X, "DECONV3X")]:
        wp = tc.pack_weights(kind, w, False)
        f0 = lambda: tc.layer_run(kind, x, None, wp, b, out_f32_nchw=True)
        f1 = lambda: tc.layer_run(kind, x, None, wp, b, out_f32_nchw=True, ref_image=img)
        f2 = lambda: tc.layer_run(kind, x, None, wp, b, out_f32_nchw=True, ref_image=img, grad_scale=0.5)
        print(f"{nm} {B}x{H}x{W}: plain {timeit(f0)*1e3:7.1f} us   +clamp+sse {timeit(f1)*1e3:7.1f} us   "
              f"+grad {timeit(f2)*1e3:7.1f} us")
        f1r = lambda: tc.layer_run(kind, x, None, wp, b, out_f32_nchw=True, ref_image=img, reduce_sse=False)
        print(f"         back-to-back: plain {timeit_b2b(f0)*1e3:7.1f} us   +clamp+sse (no reduce kernel) "
              f"{timeit_b2b(f1r)*1e3:7.1f} us")
        if "--trace" in sys.argv:
            trace(f1)
