"""Short command for ncu: the graphed tensor-core training step bench.py times (batch 8 of 3x256x256, direct form,
fused clamp + Adam appended), `iters` replays after construction (3 eager warm-ups + capture)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import iclr_17_compression_b200 as P
from iclr_17_compression_b200.train_tc import GraphedTrainStep

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.manual_seed(234)
m = P.ImageCompressor(128, return_mode="mse").cuda().train()
x = torch.rand(8, 3, 256, 256, device="cuda")
step = GraphedTrainStep(m, x, 8192.0, optimizer="adam")
torch.cuda.synchronize()
print("== replays start")
for _ in range(iters):
    loss = step(x)
torch.cuda.synchronize()
print("ok", loss.item(), step.mse.item(), step.bpp.item(), m.last_path)
