"""5x5 / 1-tap tensor-core weight gradients at the training shapes, back-to-back launches (CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iclr_17_compression_b200 import tc

g = torch.Generator().manual_seed(3)
def t(fn, n=30):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
for (B, hs) in [(8, 32), (8, 16)]:
    small = torch.randn(B, hs, hs, 128, generator=g).cuda().to(torch.bfloat16)
    big = torch.randn(B, 2 * hs, 2 * hs, 128, generator=g).cuda().to(torch.bfloat16)
    tmp = torch.zeros(5, 5, 128, 128, device="cuda")
    out = torch.zeros(128, 128, 5, 5, device="cuda")
    print(f"wgrad 5x5 {B}x{hs}x{hs}: {t(lambda: tc.wgrad(small, big, 5, 2, 2, tmp=tmp, out=out)):.1f} us (incl. untap)")
for (B, hs) in [(8, 64), (8, 32)]:
    a_ = torch.randn(B, hs, hs, 128, generator=g).cuda().to(torch.bfloat16)
    b_ = torch.randn(B, hs, hs, 128, generator=g).cuda().to(torch.bfloat16)
    tmp = torch.zeros(1, 1, 128, 128, device="cuda")
    print(f"wgrad 1tap {B}x{hs}x{hs}: {t(lambda: tc.wgrad(a_, b_, 1, 1, 0, tmp=tmp)):.1f} us")
