"""torchrun micro-benchmark of the fused NVLink all-reduce (barrier + reduce + barrier in one launch): time per call
over message size and CTA count, back-to-back launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from iclr_17_compression_b200.ddp import SymmAllReduce

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
mc = os.environ.get("MC", "1") == "1"
ar = SymmAllReduce(1768579, dev, use_multicast=mc)
ar.buf.normal_()
if rank == 0:
    print("multicast", ar.multicast, flush=True)
def timed(fn, iters=100):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
for count in (64, 65536, 442144, ar.count):
    row = []
    for blocks in (32, 74, 148, 296):
        row.append(timed(lambda: ar.all_reduce_fused(0, count, max_blocks=blocks)))
    if rank == 0:
        print(f"count {count:8d} ({count*4/1e6:6.2f} MB): " + "  ".join(f"{b} CTAs {t:5.1f} us" for b, t in zip((32, 74, 148, 296), row)), flush=True)
dist.destroy_process_group()
