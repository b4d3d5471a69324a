"""Short command for ncu: a few eval forwards of the tcgen05 pipeline at the bench shape (8 images)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import iclr_17_compression_b200 as P

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(234)
m = P.ImageCompressor(128).cuda().eval()
x = torch.rand(B, 3, 512, 768, device="cuda")
with torch.no_grad():
    for _ in range(iters):
        out = m(x)
torch.cuda.synchronize()
print("ok", out[2].item(), m.last_path)
