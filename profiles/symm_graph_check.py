"""torchrun check (N >= 2) of the graphed training step's gradient all-reduce variants: NCCL after the replay (the
reference), libb2c's fused NVLink kernel after the replay ("symm"), as the last node of the step graph ("symm_graph")
and in two buckets inside the graph ("symm_graph2").  Same weights, inputs and noise: averaged gradients must agree per tensor (atomics reorder sums slightly)
and be identical on every rank; prints the step time of each variant."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import iclr_17_compression_b200 as P
from iclr_17_compression_b200.train_tc import GraphedTrainStep

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
world = dist.get_world_size()
torch.manual_seed(234)
g = torch.Generator(device=dev).manual_seed(1000 + rank)
x = torch.rand(8, 3, 256, 256, device=dev, generator=g)
noise = torch.rand(8, 128, 16, 16, device=dev, generator=g) - 0.5
res = {}
MODES = ("nccl", "symm", "symm_graph", "symm_graph2")
for mode in MODES:
    torch.manual_seed(234)
    m = P.ImageCompressor(128).to(dev).train()
    step = GraphedTrainStep(m, x, 8192.0, allreduce=mode)
    for _ in range(3):
        step(x, noise)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        step(x, noise)
    e1.record(); torch.cuda.synchronize()
    flat = step.flat.clone()
    same = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(same, flat)
    res[mode] = (flat, all(torch.equal(same[0], t) for t in same), e0.elapsed_time(e1) / 30)
    del step, m
ok = True
for mode in MODES[1:]:
    a, b = res[mode][0].double(), res["nccl"][0].double()
    rel = ((a - b).norm() / b.norm()).item()
    if rank == 0:
        print(f"{mode}: rel-L2 vs nccl {rel:.2e}, identical on all ranks {res[mode][1]}, step {res[mode][2]:.4f} ms "
              f"(nccl {res['nccl'][2]:.4f} ms)", flush=True)
    ok &= rel < 5e-3 and res[mode][1]
assert ok
dist.destroy_process_group()
