"""ms per graphed training step (batch 8 of 3x256x256, direct form, fused clamp + Adam in the graph) - the `train`
figure of bench.py without the rest of the bench; for A/B runs of environment knobs:  python profiles/time_train.py [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import iclr_17_compression_b200 as P
from iclr_17_compression_b200.train_tc import GraphedTrainStep

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
torch.manual_seed(234)
m = P.ImageCompressor(128, return_mode="mse").cuda().train()
x = torch.rand(8, 3, 256, 256, device="cuda")
step = GraphedTrainStep(m, x, 8192.0, optimizer="adam")
step.x.copy_(x)
for _ in range(10):
    step(step.x)
torch.cuda.synchronize()
best = []
for _ in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        loss = step(step.x)
    b.record()
    torch.cuda.synchronize()
    best.append(a.elapsed_time(b) / 50)
best.sort()
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("B2C_"))
print(f"train step: min {best[0]*1e3:.1f} us  median {best[len(best)//2]*1e3:.1f} us  loss {loss.item():.3f}  [{tag}]")
