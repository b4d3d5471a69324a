"""Host->device bandwidth of all ranks copying at the same time, with the pinned buffer allocated (a) wherever the
process happens to run and (b) after binding the process to the CPUs next to its GPU (NVML cpu affinity) so that the
pinned pages are on the GPU's NUMA node.  torchrun --nproc-per-node N profiles/numa_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def h2d_gbs(buf, dev_buf, n=10):
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        dev_buf.copy_(buf, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return buf.numel() * 4 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9


dev_buf = torch.empty(24 * 3 * 512 * 768, device="cuda")
before = sorted(os.sched_getaffinity(0))
a = torch.empty(dev_buf.numel()).pin_memory()
a.uniform_()
g0 = h2d_gbs(a, dev_buf)
info = ""
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(local)
    pynvml.nvmlDeviceSetCpuAffinity(h)
    after = sorted(os.sched_getaffinity(0))
    info = f"cpus {len(before)} -> {len(after)} [{after[0]}..{after[-1]}]"
except Exception as e:      # noqa: BLE001
    info = f"nvml affinity failed: {e!r}"
b = torch.empty(dev_buf.numel()).pin_memory()
b.uniform_()
g1 = h2d_gbs(b, dev_buf)
g0b = h2d_gbs(a, dev_buf)
numa = ""
try:
    numa = open("/sys/devices/system/node/online").read().strip()
except OSError:
    pass
for r in range(world):
    barrier()
    if r == rank:
        print(f"rank {rank}: H2D {g0:.1f} GB/s unbound buffer, {g1:.1f} GB/s buffer allocated after binding "
              f"({g0b:.1f} old buffer again); {info}; numa nodes {numa}; cpu_count {os.cpu_count()}", flush=True)
if world > 1:
    dist.destroy_process_group()
