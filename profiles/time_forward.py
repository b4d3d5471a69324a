"""Device time of the eval forward at the bench shape without per-launch event records (B2C_PDL=0/1 comparison)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import iclr_17_compression_b200 as P

B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
torch.manual_seed(234)
m = P.ImageCompressor(128).cuda().eval()
x = torch.rand(B, 3, 512, 768, device="cuda")
with torch.no_grad():
    for _ in range(5):
        m(x)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            m(x)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 20)
print(f"PDL={os.environ.get('B2C_PDL', '1')} B={B} ms/step {best:.4f}  Mpix/s {B * 512 * 768 / best / 1e3:.0f}")
