"""CUPTI timeline of the data-parallel training step (graph replay + fused NVLink all-reduce) on rank 0:
    python -m torch.distributed.run --nproc-per-node N profiles/trace_train_ddp.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile
import iclr_17_compression_b200 as P
from iclr_17_compression_b200.train_tc import GraphedTrainStep

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.manual_seed(234)
m = P.ImageCompressor(128, return_mode="mse").cuda().train()
x = torch.rand(8, 3, 256, 256, device="cuda")
step = GraphedTrainStep(m, x, 8192.0)
step.x.copy_(x)
for _ in range(10):
    step(step.x)
torch.cuda.synchronize(); dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50):
    step(step.x)
b.record(); torch.cuda.synchronize()
print(f"rank {rank}: {a.elapsed_time(b) / 50 * 1e3:.1f} us per step", flush=True)
dist.barrier()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(6):
        step(step.x)
    torch.cuda.synchronize()
if rank == 0:
    prof.export_chrome_trace("/tmp/ddp_trace.json")
    ev = [e for e in json.load(open("/tmp/ddp_trace.json"))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    ev.sort(key=lambda e: e["ts"])
    starts = [i for i, e in enumerate(ev) if "pack_many" in e["name"] and e["args"].get("grid") == [128, 2, 1]]
    s0, s1 = starts[3], starts[4]
    t0 = ev[s0]["ts"]
    for e in ev[s0:s1]:
        if e["ts"] - t0 > 200 or "allreduce" in e["name"]:
            print(f"{e['ts']-t0:8.1f} {e['dur']:7.1f} s{e['args'].get('stream')} {e['name'][:60]} {e['args'].get('grid')}")
    print("step period", ev[s1]["ts"] - t0)
dist.barrier()
dist.destroy_process_group()
