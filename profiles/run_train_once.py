"""Short command for ncu: a few tensor-core training steps (batch 8 of 3x256x256), eager (no CUDA graph)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import iclr_17_compression_b200 as P

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.manual_seed(234)
m = P.ImageCompressor(128, return_mode="mse").cuda().train()
x = torch.rand(8, 3, 256, 256, device="cuda")
for _ in range(iters):
    m.zero_grad()
    _, mse, bpp = m(x)
    (8192.0 * mse + bpp).backward()
torch.cuda.synchronize()
print("ok", mse.item(), bpp.item(), m.last_path)
