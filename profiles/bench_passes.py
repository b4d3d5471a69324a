"""HBM-bound passes on bandwidth-sized tensors (at Kodak sizes they move 19-47 MB and are launch bound):
quantizer + entropy (inference / training / backward) and the sigmoid-binarizer, timed warm inside a CUDA graph of
10 back-to-back launches; algorithmic bytes per element as stated in DESIGN.md 4.3.  Prints one JSON object."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import iclr_17_compression_b200 as P
from iclr_17_compression_b200 import _lib, functional as F

lib = _lib.load()
torch.manual_seed(0)
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists("MEASURED_PEAKS.json") else 6650.0
m = P.ImageCompressor(128).cuda().eval()
params = [p.detach().reshape(-1) for p in m.bitEstimator.param_list()]
B, C, H, W = 24, 128, 96, 128            # the latents of 24 2K images (cfg5 shape): 37.7 M elements
y = (torch.randn(B, C, H, W, device="cuda") * 3).contiguous()
noise = torch.rand_like(y) - 0.5
n = y.numel()
vp = ctypes.c_void_p
def cs(): return vp(torch.cuda.current_stream().cuda_stream)
def p_(t): return vp(t.data_ptr())
z, zh = torch.empty_like(y), torch.empty((B, H, W, C), device="cuda", dtype=torch.bfloat16)
ib, pb = torch.empty(B, device="cuda"), torch.empty(B, C, device="cuda")
wsn = lib.b2c_quant_entropy_infer_workspace_bytes(B, H * W)
ws = torch.empty(wsn, device="cuda", dtype=torch.uint8)
pp, keep = _lib.ptr_array(params)
lut = F.entropy_lut(params)
dz = torch.empty_like(y)
dpar = [torch.zeros_like(p) for p in params]
dp, keep2 = _lib.ptr_array(dpar)
gb = torch.ones(1, device="cuda")
sig, code = torch.empty_like(y), torch.empty_like(y)
cases = {
    "quant_entropy_infer (read y, write z fp32 + z bf16)": (10, lambda: lib.b2c_quant_entropy_infer(p_(y), pp, p_(z), p_(zh), p_(ib), vp(0), B, C, H * W, p_(ws), wsn, cs())),
    "quant_entropy_infer_lut (read y, write z fp32 + z bf16, table)": (10, lambda: lib.b2c_quant_entropy_infer_lut(p_(y), pp, p_(lut), 32, p_(z), p_(zh), p_(ib), vp(0), B, C, H * W, p_(ws), wsn, cs())),
    "quant_entropy_fwd train (read y, noise, write z)": (12, lambda: lib.b2c_quant_entropy_fwd(p_(y), p_(noise), pp, p_(z), p_(pb), p_(ib), B, C, H * W, cs())),
    "entropy_bwd (read z, write dz)": (8, lambda: lib.b2c_entropy_bwd(p_(z), pp, p_(gb), p_(dz), dp, B, C, H * W, cs())),
    "sigmoid_binarize_fwd (read y, write sig, code, code bf16)": (14, lambda: lib.b2c_sigmoid_binarize_fwd(p_(y), p_(sig), p_(code), p_(zh), B, C, H * W, cs())),
    "copy (torch, read + write fp32)": (8, lambda: z.copy_(y)),
}
only = os.environ.get("B2C_PASSES_ONLY")
if only:
    cases = {k: v for k, v in cases.items() if only in k}
out = {"elements": n, "hbm_peak_gbs": peak}
R = 10
for name, (bpe, fn) in cases.items():
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(R):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / R
    gbs = bpe * n / (ms * 1e-3) / 1e9
    out[name] = {"ms": round(ms, 4), "bytes_per_element": bpe, "GB/s": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peak, 3),
                 "Gelem/s": round(n / ms / 1e6, 1)}
print(json.dumps(out, indent=1))
