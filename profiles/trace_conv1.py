"""Debug: clock64 timeline of CTA 0 of the persistent CONV1+GDN1 kernel (b2c_tc_debug_set_trace).
Worker marks per steady-state iteration: patch, kb0+1 built, acc ready, x2[0], kb2 + x2[1], kb3, norm ready, stored."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import iclr_17_compression_b200 as P
from iclr_17_compression_b200 import _lib, tc

lib = _lib.load()
lib.b2c_tc_debug_set_trace.argtypes = [ctypes.c_void_p]
torch.manual_seed(234)
m = P.ImageCompressor(128).cuda().eval()
x = torch.rand(24, 3, 512, 768, device="cuda")
Pk = m._packed_weights()
E = m.Encoder
fn = lambda: tc.layer_run(tc.TC_CONV1, x, None, Pk["conv1"], E.conv1.bias, gdn=Pk["gdn1"], gdn_mode=1, split3=True)
with torch.no_grad():
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print("conv1+gdn1 ms", e0.elapsed_time(e1) / 20)
    buf = torch.zeros(4 * 64, dtype=torch.int64, device="cuda")
    lib.b2c_tc_debug_set_trace(ctypes.c_void_p(buf.data_ptr()))
    fn()
    torch.cuda.synchronize()
    lib.b2c_tc_debug_set_trace(ctypes.c_void_p(0))
t = buf.cpu().view(4, 64)
t0 = min(int(v) for v in t.flatten() if v > 0)
for role, nm in enumerate(["producer", "mma", "worker / epilogue", "builder"]):
    vals = [int(v) - t0 for v in t[role] if v > 0]
    if not vals:
        continue
    print(nm, vals)
    print("   deltas", [b - a for a, b in zip(vals, vals[1:])])
