#!/usr/bin/env python
"""Benchmark of the Balle-2017 codec hot path (BASELINE.json metric: Mpix/s encode+decode at 768x512).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A step = one ImageCompressor eval forward (encode -> round -> bits -> decode -> clamp/SSE) over a batch
of 24 synthetic Kodak-shape images (BASELINE.json configs[1]) per GPU.  Multi-GPU inference shards per
image with no collective (weak scaling: 24 images per rank).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

IMAGES, H, W, N_CH = 24, 512, 768, 128
MPIX_PER_IMAGE = H * W / 1e6
FLOP_PER_PIXEL = 45015.0          # convs 39776 + GDN matmuls 5120 + pointwise 119
METRIC = "Mpix/s encode+decode (768x512) per B200"
E2E_CHUNK = int(os.environ.get("B2C_E2E_CHUNK", "3"))


def workload_config(world):
    """the `config` object of the JSON line - the SAME for both arms (the reference arm times a bounded sample of this
    workload; what the sample was is in its cpu_baseline.sample)"""
    return {"workload": f"kodak-shape inference, {IMAGES}x3x{H}x{W} per GPU, N={N_CH}, eval, one batch "
                        f"of {IMAGES} (configs[1])",
            "weights": "seed-234 random init", "parallelism": f"per-image sharding x{world}, no collective",
            "l2": "per-step working set (226 MB in/out + >600 MB intermediates) exceeds the 126 MB L2"}


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled under the load of the timed region (B200_PROFILING.md).  The K
    timed steps last only a few ms and nvidia-smi answers every ~50-100 ms (later still when 8 ranks start at once), so
    the sampler is started BEFORE the warm-up, every row is time-stamped, and the caller keeps the SAME step loop
    running right after the timed steps until rows have arrived; only rows inside that load window are reported."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def count_since(self, t0):
        return sum(1 for t, _ in list(self.rows) if t >= t0)

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, r in self.rows:
            try:
                smax = float(r[1])
                if (t0 is not None and t < t0) or (t1 is not None and t > t1):
                    continue
                sm.append(float(r[0]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "window": "timed steps + the same step loop kept running until >= 3 samples"}


# ------------------------------------------------------------------------------------------ reference / CPU arm
def cpu_reference_throughput(budget_s, images_per_step, steps=None, warmup=1):
    """Times the reference algorithm's CPU path (the oracle port: plain torch CPU ops, all host threads) on a
    bounded sample of the same workload.  Returns (Mpix/s, cores, sample description, ms per step)."""
    from oracle import balle17_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = O.init_state_dict(N_CH, seed=234)
    x = O.synth_images("U", images_per_step, H, W, 1002)
    with torch.no_grad():
        for _ in range(warmup):
            O.forward(x[:1], sd)
        t0 = time.perf_counter()
        done = 0
        nsteps = 0
        while True:
            for i in range(images_per_step):          # batch 1 per call, as the reference's testKodak does (train.py:162)
                O.forward(x[i:i + 1], sd)
                done += 1
            nsteps += 1
            el = time.perf_counter() - t0
            if (steps is not None and nsteps >= steps) or (steps is None and el >= budget_s):
                break
    return (done * MPIX_PER_IMAGE / el, cores,
            f"{done} eval forwards of single 3x{H}x{W} images (batch 1 as train.py:162 does, {images_per_step} images per "
            f"step, cycling through {images_per_step} of the {IMAGES} images)", el / nsteps * 1e3)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = 2
    val, cores, sample, ms = cpu_reference_throughput(None, per_step, steps=args.steps, warmup=max(1, args.warmup))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Mpix/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": val, "unit": "Mpix/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ B200 arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--no-train", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    # exactly ONE line may reach stdout: libraries (NCCL prints its version banner there) write to stderr instead
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    import torch.distributed as dist
    import iclr_17_compression_b200 as P
    from iclr_17_compression_b200 import _lib, functional as B2F
    from oracle import balle17_oracle as O      # synthetic inputs + cpu_baseline leg only

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    torch.manual_seed(234)
    model = P.ImageCompressor(N_CH).to(dev).eval()
    x_host = O.synth_images("U", IMAGES, H, W, 1002 + rank).pin_memory()
    x = x_host.to(dev)

    def step(inp):
        with torch.no_grad():
            return model(inp)

    sampler = ClockSampler(local)
    if rank == 0:                                 # the line reports rank 0's GPU; 8 concurrent nvidia-smi loops only slow each other
        sampler.start()                           # before the warm-up: nvidia-smi is answering by the time the load starts
    for _ in range(args.warmup):
        step(x)
    barrier()

    # ---- timed region: K steps, device time, inputs resident in HBM.  No per-launch event records in here: an
    # event between two kernels defeats the programmatic dependent launch of the next layer (and costs ~1 us each).
    n0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_load0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        step(x)
    ev1.record()
    barrier()
    launches = _lib.launch_count() - n0
    ms = ev0.elapsed_time(ev1) / args.steps
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    value = IMAGES * world * MPIX_PER_IMAGE / (ms / 1e3)

    # ---- per-kernel pass: the same K steps again with a CUDA-event pair around every libb2c launch (on the
    # launching stream) -> the dominant kernel's average duration and its share of the step for the roofline
    timer = B2F.start_timing()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record()
    for _ in range(args.steps):
        step(x)
    ev3.record()
    barrier()
    B2F.stop_timing()
    while (sampler.proc is not None and sampler.proc.poll() is None and sampler.count_since(t_load0) < 3
           and time.perf_counter() - t_load0 < 8.0):
        for _ in range(args.steps):               # the same load, untimed, until the sampler has seen it
            step(x)
        torch.cuda.synchronize()
    clocks = sampler.stop(t_load0, time.perf_counter())
    ms_instrumented = ev2.elapsed_time(ev3) / args.steps

    # ---- end to end: pinned host input -> H2D -> forward -> D2H of latents + per-image (bits, sse, sse_clipped),
    # through the package's host-facing call (chunks of 4 images, copies overlapped with the previous chunk's kernels)
    from iclr_17_compression_b200.inference import HostPipeline
    pipe = HostPipeline(model, IMAGES, H, W, chunk=E2E_CHUNK, device=dev)
    for _ in range(2):
        pipe.compress(x_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pipe.compress(x_host)
    barrier()
    e2e_ms = (time.perf_counter() - t0) / args.steps * 1e3
    t = torch.tensor([e2e_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = t.item()
    e2e_val = IMAGES * world * MPIX_PER_IMAGE / (e2e_ms / 1e3)
    # the same call fed with the 8-bit images the fp32 tensors are made from (ToTensor's / 255 done on the device):
    # reported next to e2e, not instead of it - the reference's API takes fp32 tensors
    u8_host = (x_host * 255).round().to(torch.uint8).pin_memory()
    pipe_u8 = HostPipeline(model, IMAGES, H, W, chunk=int(os.environ.get("B2C_E2E_U8_CHUNK", "12")), device=dev)
    for _ in range(2):
        pipe_u8.compress_u8(u8_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pipe_u8.compress_u8(u8_host)
    barrier()
    u8_ms = (time.perf_counter() - t0) / args.steps * 1e3
    t = torch.tensor([u8_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    u8_ms = t.item()
    # the bound of that call: one pinned-host -> device copy of the step's input alone (PCIe), same CUDA-event timing
    xd = torch.empty_like(x)
    xd.copy_(x_host, non_blocking=True)
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(3):
        xd.copy_(x_host, non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    h2d_ms = c0.elapsed_time(c1) / 3
    del xd

    # ---- roofline of the dominant kernel (largest share of the step), from the in-region event pairs
    pk, pk_kind = peaks()
    ops = timer.summary()
    total_ms = sum(d["ms_total"] for d in ops.values())
    top_name, top = max(ops.items(), key=lambda kv: kv[1]["ms_total"])
    tensor_bound = top_name.startswith("conv") or top_name.startswith("tc_layer")
    if tensor_bound:
        achieved = top["flops"] / (top["ms_avg"] * 1e-3) / 1e12
        peak = pk["bf16_tflops_sustained"]
        roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak}
        if "split1" in top_name:
            # the analysis transform runs error-compensated bf16x3 (hi*hi + hi*lo + lo*hi): 3 MMAs per algorithmic
            # product, needed for bit-exact rounded latents - the tensor pipe executes 3x `achieved`
            roof["executed_tflops"] = 3 * achieved
            roof["executed_frac_of_peak"] = 3 * achieved / peak
            roof["executed_frac_of_burst_peak"] = 3 * achieved / pk.get("bf16_tflops", peak)
    else:
        achieved = top["bytes"] / (top["ms_avg"] * 1e-3) / 1e9
        peak = pk["hbm_gbs"]
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak}
    traffic = None
    try:   # DRAM bytes of this kernel from the committed ncu --set full capture of the same batch (24 images)
        nt = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")))
        for key, v in nt["kernels"].items():
            if top_name.startswith(key):
                traffic = v["dram_mbytes_per_launch"] * 1e6 * IMAGES / nt["images"]
                roof["tensor_pipe_active_pct_ncu"] = v["tensor_pipe_active_pct"]
    except Exception:
        pass
    roof.update({"traffic": traffic, "traffic_unit": "bytes/launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, "
                                                     "profiles/r2_ncu_traffic.json, same batch)",
                 "kernel": top_name, "kernel_ms": top["ms_avg"],
                 "share_of_step": top["ms_total"] / total_ms, "peak_source": pk_kind,
                 "kernel_timing": "CUDA-event pair around every launch in a second pass of the same K steps "
                                  f"({ms_instrumented:.4f} ms/step with the event records, {ms:.4f} without)",
                 "pipeline_frac_of_tensor_roofline": (FLOP_PER_PIXEL * IMAGES * H * W / (ms * 1e-3) / 1e12)
                 / pk["bf16_tflops_sustained"]})

    line = {"metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world),
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_val, "unit": "Mpix/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(pipe.h2d_bytes), "d2h_bytes_per_step": int(pipe.d2h_bytes),
                    "call": f"inference.HostPipeline.compress (slices of {pipe.sizes} images, copy/compute overlap)",
                    "h2d_copy_alone_ms": h2d_ms, "h2d_copy_alone_gbs": x.numel() * 4 / (h2d_ms * 1e-3) / 1e9,
                    "bound": "PCIe host->device copy of the fp32 input"},
            "e2e_u8_input": {"value": IMAGES * world * MPIX_PER_IMAGE / (u8_ms / 1e3), "unit": "Mpix/s", "ms_per_step": u8_ms,
                             "h2d_bytes_per_step": int(pipe.h2d_bytes // 4), "d2h_bytes_per_step": int(pipe.d2h_bytes),
                             "call": f"inference.HostPipeline.compress_u8 (uint8 images from pinned host memory, slices of "
                                     f"{pipe_u8.sizes}, x / 255 on the device; identical results)"},
            "roofline": roof,
            "ops": {k: {"ms_avg": round(v["ms_avg"], 4), "calls": v["calls"]} for k, v in ops.items()}}

    # ---- training step (cfg3): batch 8 of 3x256x256, noise quantizer, fwd+bwd of 8192*mse + bpp
    if not args.no_train:
        try:
            line["train"] = train_bench(model, dev, world, rank, barrier)
        except Exception as e:   # the inference line must still print
            line["train"] = {"error": repr(e)}

    if rank == 0 and world == 1:
        val, cores, sample, _ = cpu_reference_throughput(args.cpu_budget, 4)
        line["cpu_baseline"] = {"value": val, "unit": "Mpix/s", "cores": cores, "kind": "port", "sample": sample}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def train_bench(model, dev, world, rank, barrier, steps=10, warmup=3):
    import torch.distributed as dist
    from oracle import balle17_oracle as O
    model.train()
    model.return_mode = "mse"
    sync = None
    if world > 1:
        from iclr_17_compression_b200.ddp import GradSync
        sync = GradSync(model)
    xb = O.synth_images("U", 8, 256, 256, 1004 + rank).to(dev)

    graphed = None
    if os.environ.get("B2C_TRAIN_GRAPH", "1") == "1" and getattr(model, "fast_training", False):
        from iclr_17_compression_b200.train_tc import GraphedTrainStep
        if sync is not None:
            sync.remove()                                      # hooks do not fire inside a graph replay;
            sync = None                                        # the graphed step all-reduces its flat gradient buffer
        graphed = GraphedTrainStep(model, xb, 8192.0, direct=os.environ.get("B2C_TRAIN_DIRECT", "1") == "1")

    if graphed is not None:
        graphed.x.copy_(xb)            # the batch sits in the step's static input buffer (where a loader's H2D copy
                                       # would put it): resident in HBM when the timed region starts, no copy per step

    def tstep():
        if graphed is not None:
            graphed(graphed.x)
            return
        for p in model.parameters():
            p.grad = None
        _, mse, bpp = model(xb)
        (8192.0 * mse + bpp).backward()
        if sync is not None:
            sync.finish()

    from iclr_17_compression_b200 import functional as B2F
    for _ in range(warmup):
        tstep()
    barrier()
    timer = B2F.start_timing()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        tstep()
    e1.record()
    barrier()
    B2F.stop_timing()
    ops = timer.summary()
    ms = e0.elapsed_time(e1) / steps
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    model.eval()
    model.return_mode = "latent"
    return {"metric": "train img/s (batch 8 of 3x256x256 per GPU, fwd+bwd, noise quantizer, lambda=8192)",
            "value": 8 * world / (ms / 1e3), "unit": "img/s", "ms_per_step": ms,
            "allreduce": ("none (1 GPU)" if world == 1 else
                          ("libb2c NVLink kernel over symmetric memory (" + ("NVLS multimem" if graphed._symm.multicast else "peer loads/stores")
                           + "), cross-rank barrier + reduction + barrier in ONE launch, "
                           + (("captured as the last node of the step graph" if getattr(graphed, "_symm_buckets", 1) == 1 else
                               "two buckets captured inside the step graph")
                              if getattr(graphed, "_symm_in_graph", False) else "after the graphed step")) if getattr(graphed, "_symm", None) is not None else
                          ("two-bucket NCCL all-reduce captured inside the step graph" if getattr(graphed, "ar_in_graph", False)
                           else "one NCCL all-reduce (mean) of the flat 7 MB gradient buffer after the graphed step")
                          if graphed is not None else
                          "nccl bucketed, overlapped with backward"),
            "input": ("batch resident in the step's static input buffer, where a loader's host-to-device copy would put it "
                      "(no per-step device-to-device copy)") if graphed is not None else "device tensor",
            "path": getattr(model, "last_path", None), "cuda_graph": graphed is not None,
            "direct": bool(getattr(graphed, "direct", False)),
            "ops": {k: round(v["ms_total"] / steps, 4) for k, v in sorted(ops.items(), key=lambda kv: -kv[1]["ms_total"])},
            "flop_per_step": 68.7e9, "frac_of_tensor_roofline": 68.7e9 / (ms * 1e-3) / 1e12 / peaks()[0]["bf16_tflops_sustained"]}


if __name__ == "__main__":
    main()
