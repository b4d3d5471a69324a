"""CUPTI timeline of one HostPipeline.compress call (the e2e figure of bench.py): memcpys and kernels."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
import iclr_17_compression_b200 as P
from iclr_17_compression_b200.inference import HostPipeline

torch.manual_seed(234)
m = P.ImageCompressor(128).cuda().eval()
x = torch.rand(24, 3, 512, 768).pin_memory()
sizes = [int(v) for v in os.environ.get("SIZES", "1,3,3,3,3,3,3,3,2").split(",")]
pipe = HostPipeline(m, 24, 512, 768, chunk=max(sizes), sizes=sizes)
for _ in range(5):
    pipe.compress(x)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(20):
    pipe.compress(x)
print("host-clock ms per call", (time.perf_counter() - t0) / 20 * 1e3)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    pipe.compress(x)
    torch.cuda.synchronize()
prof.export_chrome_trace("/tmp/e2e.json")
ev = [e for e in json.load(open("/tmp/e2e.json"))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
for e in ev:
    if e["cat"] == "gpu_memcpy" or "deconv3x" in e["name"] or "conv1" in e["name"]:
        print(f"{e['ts']-t0:9.1f} {e['dur']:8.1f} s{e['args'].get('stream')} {e['name'][:50]} {e['args'].get('bytes', '')}")
print("span", ev[-1]["ts"] + ev[-1]["dur"] - t0)
