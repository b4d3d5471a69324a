"""Measures the BASELINE.json configurations that bench.py's single headline line does not cover, plus the
reference algorithm executed by PyTorch eager (cuDNN) on the same B200 as a second comparator
(BASELINE.md section 4, item 3).  Writes profiles/r2_configs.json.

    python profiles/bench_configs.py            # on a B200 box
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import iclr_17_compression_b200 as P
from iclr_17_compression_b200.inference import HostPipeline
from oracle import balle17_oracle as O          # comparator + synthetic inputs (test infrastructure)


def timed(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


out = {"gpu": torch.cuda.get_device_name(0)}
torch.manual_seed(234)
m = P.ImageCompressor(128).cuda().eval()
sd = {k: v.detach() for k, v in m.state_dict().items()}

# ---- cfg2: 24 Kodak-shape images, (i) batch 1 x 24 as the reference's testKodak does, (ii) one batch of 24
x24 = O.synth_images("U", 24, 512, 768, 1002).cuda()
mpix24 = 24 * 512 * 768 / 1e6
with torch.no_grad():
    ms_b24 = timed(lambda: m(x24), 20)
    ms_b1 = timed(lambda: [m(x24[i:i + 1]) for i in range(24)], 5)
out["cfg2_batch24"] = {"ms": ms_b24, "mpix_s": mpix24 / ms_b24 * 1e3}
out["cfg2_batch1_x24_eager_launches"] = {"ms": ms_b1, "mpix_s": mpix24 / ms_b1 * 1e3,
                                         "note": "plain model(x[i:i+1]) per image: one CUDA-graph replay per call from the second call on "
                                                 "(eager: host-launch bound, ~14 launches + TMA descriptors per image, 9.7 ms)"}
pipe = HostPipeline(m, 24, 512, 768, chunk=1)
xh = x24.cpu().pin_memory()
pipe.compress(xh)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    pipe.compress(xh)
ms = (time.perf_counter() - t0) / 5 * 1e3
out["cfg2_batch1_x24_host_pipeline_graphs"] = {"ms": ms, "mpix_s": mpix24 / ms * 1e3,
                                               "note": "HostPipeline chunk=1: CUDA graph per image, H2D/D2H overlapped"}

# ---- reference algorithm on the same GPU: PyTorch eager / cuDNN (what users of the reference get today)
for tf32 in (True, False):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    with torch.no_grad():
        ms_ref = timed(lambda: O.forward(x24, sd), 5)
        ms_ref1 = timed(lambda: [O.forward(x24[i:i + 1], sd) for i in range(24)], 3)
    out[f"cfg2_reference_eager_cuda_tf32_{int(tf32)}"] = {"batch24_ms": ms_ref, "batch24_mpix_s": mpix24 / ms_ref * 1e3,
                                                          "batch1_x24_ms": ms_ref1,
                                                          "batch1_x24_mpix_s": mpix24 / ms_ref1 * 1e3}
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

# ---- cfg3: training step, batch 8 of 3x256x256 (eager fp32 kernels, eager bf16 tensor-core, graphed)
xb = O.synth_images("U", 8, 256, 256, 1004).cuda()
mt = P.ImageCompressor(128, return_mode="mse").cuda().train()


def step():
    for p in mt.parameters():
        p.grad = None
    _, mse, bpp = mt(xb)
    (8192.0 * mse + bpp).backward()


mt.fast_training = False
out["cfg3_train_fp32_kernels"] = {"ms": timed(step, 10), }
mt.fast_training = True
out["cfg3_train_tc_bf16_eager"] = {"ms": timed(step, 10)}
from iclr_17_compression_b200.train_tc import GraphedTrainStep
gs = GraphedTrainStep(mt, xb, 8192.0)
out["cfg3_train_tc_bf16_graphed"] = {"ms": timed(lambda: gs(xb), 20)}
del gs
# whole iteration (forward + backward + gradient clamp + Adam) as one graph replay, and larger per-GPU batches: the
# batch-8 step is a chain of ~20 dependent launches on <= 256 CTAs (latency bound); these show the same kernels with
# enough tiles to fill the machine
mo = P.ImageCompressor(128, return_mode="mse").cuda().train()
go = GraphedTrainStep(mo, xb, 8192.0, optimizer="adam")
out["cfg3_train_tc_bf16_graphed_with_clamp_adam"] = {"ms": timed(lambda: go(xb), 20)}
del go, mo
for bsz in (32, 64):
    xbig = O.synth_images("U", bsz, 256, 256, 1004).cuda()
    mb = P.ImageCompressor(128, return_mode="mse").cuda().train()
    gb = GraphedTrainStep(mb, xbig, 8192.0)
    ms_big = timed(lambda: gb(xbig), 10)
    out[f"train_tc_bf16_graphed_batch{bsz}"] = {"ms": ms_big, "img_s": bsz / ms_big * 1e3,
                                                "frac_of_tensor_roofline": 8.59e9 * bsz / (ms_big * 1e-3) / 1e12 / 1396.4}
    del gb, mb, xbig
leaf = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
noise = O.synth_noise(8, 128, 256, 256, 2002).cuda()


def ref_step():
    for v in leaf.values():
        v.grad = None
    o = O.forward(xb, leaf, training=True, noise=noise)
    (8192.0 * o["mse"] + o["bpp"]).backward()


torch.backends.cudnn.allow_tf32 = True
out["cfg3_reference_eager_cuda_tf32_1"] = {"ms": timed(ref_step, 5)}
torch.backends.cudnn.allow_tf32 = False
out["cfg3_reference_eager_cuda_tf32_0"] = {"ms": timed(ref_step, 5)}
for k in list(out):
    if k.startswith("cfg3") and "ms" in out[k]:
        out[k]["img_s"] = 8 / out[k]["ms"] * 1e3

# ---- cfg5 stand-in: 2K images (3x1536x2048), ImageCompressor(128) on tensor cores and (192) on the fp32 kernels
x2k = O.synth_images("S", 4, 1536, 2048, 1003).cuda()
with torch.no_grad():
    ms2k = timed(lambda: m(x2k), 5)
out["cfg5_2k_n128_tc_batch4"] = {"ms": ms2k, "mpix_s": 4 * 1536 * 2048 / 1e6 / ms2k * 1e3}
for N in (192, 256):
    mw = P.ImageCompressor(N).cuda().eval()
    sdw = {k: v.detach() for k, v in mw.state_dict().items()}
    with torch.no_grad():
        ms_tc = timed(lambda: mw(x2k), 5)
        mw.fast_inference = False
        ms_fp32 = timed(lambda: mw(x2k[:1]), 2, warmup=1)
        mw.fast_inference = True
        torch.backends.cudnn.allow_tf32 = True
        ms_ref = timed(lambda: O.forward(x2k[:1], sdw), 2, warmup=1)
        torch.backends.cudnn.allow_tf32 = False
    out[f"cfg5_2k_n{N}_tc_batch4"] = {"ms": ms_tc, "mpix_s": 4 * 1536 * 2048 / 1e6 / ms_tc * 1e3}
    out[f"cfg5_2k_n{N}_fp32_kernels_batch1"] = {"ms": ms_fp32, "mpix_s": 1536 * 2048 / 1e6 / ms_fp32 * 1e3}
    out[f"cfg5_2k_n{N}_reference_eager_cuda_tf32_batch1"] = {"ms": ms_ref, "mpix_s": 1536 * 2048 / 1e6 / ms_ref * 1e3}
    del mw

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", "r2_configs.json")
os.makedirs(os.path.dirname(path), exist_ok=True)
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
