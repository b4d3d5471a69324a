"""Timeline of one graphed training step (kernel name, stream, start, duration) from torch.profiler's CUPTI activity
records: which kernels are on the critical path, where the gaps are.  python profiles/trace_train.py [out.json]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import iclr_17_compression_b200 as P
from iclr_17_compression_b200.train_tc import GraphedTrainStep

torch.manual_seed(234)
m = P.ImageCompressor(128, return_mode="mse").cuda().train()
x = torch.rand(8, 3, 256, 256, device="cuda")
step = GraphedTrainStep(m, x, 8192.0, optimizer="adam")
step.x.copy_(x)
for _ in range(10):
    step(step.x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(6):
        step(step.x)
    torch.cuda.synchronize()
path = "/tmp/train_trace.json"
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
# the 5th step: from the 5th occurrence of the first-layer packing kernel to the next one
starts = [i for i, e in enumerate(ev) if "pack_many" in e["name"] and e["args"].get("grid") == [128, 2, 1]]
a, b = starts[4], starts[5]
t0 = ev[a]["ts"]
rows = []
for e in ev[a:b]:
    rows.append({"t_us": round(e["ts"] - t0, 2), "dur_us": round(e["dur"], 2), "stream": e["args"].get("stream"),
                 "name": e["name"][:70], "grid": e["args"].get("grid"), "block": e["args"].get("block")})
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/train_timeline.json"
json.dump(rows, open(out, "w"), indent=0)
end = max(r["t_us"] + r["dur_us"] for r in rows)
print(f"step span {end:.1f} us, {len(rows)} activities")
for r in rows:
    print(f"{r['t_us']:8.1f} {r['dur_us']:7.1f}  s{r['stream']:<4} {r['name'][:60]:60s} {r['grid']}")
