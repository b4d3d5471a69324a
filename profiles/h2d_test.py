"""Pinned host -> device copy of the bench input (113 MB fp32) with 1, 2 and 4 concurrent streams."""
import time, torch
x = torch.rand(24, 3, 512, 768).pin_memory()
d = torch.empty_like(x, device="cuda")
for ns in (1, 2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(ns)]
    parts = list(zip(x.chunk(ns), d.chunk(ns)))
    def run():
        for s, (a, b) in zip(streams, parts):
            with torch.cuda.stream(s):
                b.copy_(a, non_blocking=True)
    for _ in range(3): run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): run()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 10 * 1e3
    print(f"{ns} streams: {ms:.3f} ms  {x.numel() * 4 / ms / 1e6:.1f} GB/s")
