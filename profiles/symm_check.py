"""torchrun check (N >= 2): ddp.SymmAllReduce (libb2c NVLink / NVLS kernel over symmetric memory) vs NCCL all-reduce:
values, and time per call on the 7 MB gradient buffer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from iclr_17_compression_b200.ddp import SymmAllReduce

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
world = dist.get_world_size()
n = 1768579
for mc in (True, False):
    ar = SymmAllReduce(n, dev, use_multicast=mc)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    src = torch.randn(ar.count, device=dev, generator=g)
    ref = src.clone()
    dist.all_reduce(ref, op=dist.ReduceOp.AVG)
    ar.buf.copy_(src)
    ar.all_reduce()
    torch.cuda.synchronize()
    err = (ar.buf - ref).abs().max().item()
    same = [torch.empty_like(ar.buf) for _ in range(world)]
    dist.all_gather(same, ar.buf)
    ident = all(torch.equal(same[0], t) for t in same)
    def timed(fn, iters=50):
        for _ in range(5): fn()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e3
    t_symm = timed(ar.all_reduce)
    t_nccl = timed(lambda: dist.all_reduce(ref, op=dist.ReduceOp.AVG))
    if rank == 0:
        print(f"requested multicast={mc} used multicast={ar.multicast}: max |symm - nccl| = {err:.3e}, identical on all ranks: {ident}, "
              f"symm {t_symm:.1f} us / call, nccl {t_nccl:.1f} us / call", flush=True)
        assert err < 1e-5 and ident
dist.destroy_process_group()
