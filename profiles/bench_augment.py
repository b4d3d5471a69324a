"""Input transform timings on one B200: b2c_resized_crop_flip on resident uint8 batches (CUDA events, L2
flushed between iterations), the same through the public call from host memory (pack + pinned copy + launch), and the
reference's own CPU transform (torchvision Compose on PIL images, datasets.py:24-34, one core = its num_workers) on the
same images and draws.  python profiles/bench_augment.py > gpurun_out/augment.json"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from iclr_17_compression_b200 import datasets as D
from oracle import resample_oracle as R

PEAK = 6543.4
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timed(fn, iters=20, warm=3):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for i in range(warm + iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


res = {"hbm_peak_gbs": PEAK}
for name, B, (h, w) in (("8x375x500", 8, (375, 500)), ("64x375x500", 64, (375, 500)), ("64x1356x2040", 64, (1356, 2040)),
                        ("256x1356x2040", 256, (1356, 2040))):
    base = [R.synth_photo(h, w, s) for s in range(4)]
    imgs = [base[n % 4] for n in range(B)]
    torch.manual_seed(5)
    params = [D.sample_params(h, w) for _ in range(B)]
    buf, desc, mh, mw = D.pack_images(imgs, params, crop_only=False)
    src, dd = buf.cuda(), desc.cuda()
    y = D.resized_crop_flip_packed(src, dd, 256, mh, mw)
    assert torch.equal(y[0].cpu(), R.resized_crop_flip(imgs[0], *params[0]))
    t = timed(lambda: D.resized_crop_flip_packed(src, dd, 256, mh, mw, check=False))
    crop_bytes = sum(p[2] * p[3] * 3 for p in params)
    out_bytes = B * 3 * 256 * 256 * 4
    t0 = time.perf_counter()
    for _ in range(3):
        D.resized_crop_flip(imgs, params, 256)
    torch.cuda.synchronize()
    t_api = (time.perf_counter() - t0) / 3 * 1e3
    r = {"images": B, "kernel_ms": t, "kernel_img_s": B / t * 1e3, "algorithmic_bytes": crop_bytes + out_bytes,
         "achieved_gbs": (crop_bytes + out_bytes) / t / 1e6, "frac_of_hbm_peak": (crop_bytes + out_bytes) / t / 1e6 / PEAK,
         "source_bytes_resident": int(buf.numel()), "public_call_from_host_ms": t_api, "public_call_img_s": B / t_api * 1e3,
         "mean_crop": [float(np.mean([p[2] for p in params])), float(np.mean([p[3] for p in params]))]}
    try:
        from PIL import Image
        from torchvision.transforms import functional as F
        pil = [Image.fromarray(i, "RGB") for i in imgs[:16]]
        t0 = time.perf_counter()
        for im, p in zip(pil, params):
            o = F.resized_crop(im, p[0], p[1], p[2], p[3], [256, 256])
            o = F.hflip(o) if p[4] else o
            o = F.vflip(o) if p[5] else o
            F.to_tensor(o)
        dt = time.perf_counter() - t0
        r["reference_cpu_1core_img_s"] = len(pil) / dt
    except ImportError as e:
        r["reference_cpu_1core_img_s"] = f"unavailable: {e}"
    res[name] = r
print(json.dumps(res, indent=1))
