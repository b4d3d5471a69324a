"""e2e (pinned host fp32 -> latents + statistics on the host) for different slice plans of inference.HostPipeline."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import iclr_17_compression_b200 as P
from iclr_17_compression_b200.inference import HostPipeline

torch.manual_seed(234)
m = P.ImageCompressor(128).cuda().eval()
x_host = torch.rand(24, 3, 512, 768).pin_memory()
xd = torch.empty(24, 3, 512, 768, device="cuda")
for _ in range(3):
    xd.copy_(x_host, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    xd.copy_(x_host, non_blocking=True)
torch.cuda.synchronize()
print("H2D copy alone ms", (time.perf_counter() - t0) / 10 * 1e3)
for plan in sys.argv[1:] or ["3", "2", "4", "6", "1"]:
    if "," in plan:
        sizes = [int(v) for v in plan.split(",")]
        pipe = HostPipeline(m, 24, 512, 768, chunk=max(sizes), sizes=sizes)
    else:
        pipe = HostPipeline(m, 24, 512, 768, chunk=int(plan))
    for _ in range(3):
        pipe.compress(x_host)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        pipe.compress(x_host)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 20 * 1e3
    print(f"plan {plan:24s} sizes {pipe.sizes}: {ms:.3f} ms  {24 * 0.393216 / ms:.2f} Gpix/s", flush=True)
