import os, sys
sys.path.insert(0, "/root/repo")
import torch
import iclr_17_compression_b200 as P
from iclr_17_compression_b200 import functional as B2F
for N in (192, 256):
    torch.manual_seed(234)
    m = P.ImageCompressor(N).cuda().eval()
    x = torch.rand(4, 3, 1536, 2048, device="cuda")
    with torch.no_grad():
        for _ in range(3): m(x)
        torch.cuda.synchronize()
        t = B2F.start_timing()
        for _ in range(5): m(x)
        torch.cuda.synchronize()
        B2F.stop_timing()
    print("N", N, m.last_path)
    for k, v in sorted(t.summary().items(), key=lambda kv: -kv[1]["ms_total"]):
        print(f"  {k:50s} {v['ms_total']/5:8.3f} ms  {v.get('tflops', 0):8.1f}")
