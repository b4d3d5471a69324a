"""Debug: clock64 timeline of CTA (0,0,0) of one tensor-core layer (b2c_tc_debug_set_trace)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import iclr_17_compression_b200 as P
from iclr_17_compression_b200 import _lib, tc

lib = _lib.load()
lib.b2c_tc_debug_set_trace.argtypes = [ctypes.c_void_p]
torch.manual_seed(234)
m = P.ImageCompressor(128).cuda().eval()
B = 24
x = torch.rand(B, 3, 512, 768, device="cuda")
Pk = m._packed_weights()
E, D = m.Encoder, m.Decoder

def run(name, fn):
    buf = torch.zeros(3 * 64, dtype=torch.int64, device="cuda")
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    lib.b2c_tc_debug_set_trace(ctypes.c_void_p(buf.data_ptr()))
    out = fn()
    torch.cuda.synchronize()
    lib.b2c_tc_debug_set_trace(ctypes.c_void_p(0))
    t = buf.cpu().view(3, 64)
    t0 = min(int(v) for v in t.flatten() if v > 0)
    print("==", name)
    for role, nm in enumerate(["producer", "mma", "worker"]):
        vals = [int(v) - t0 for v in t[role] if v > 0]
        print(f"  {nm:8s}", vals)
    return out

with torch.no_grad():
    h1 = run("conv1+gdn1", lambda: tc.layer_run(tc.TC_CONV1, x, None, Pk["conv1"], E.conv1.bias, gdn=Pk["gdn1"], gdn_mode=1, split3=True))
    h2 = run("conv2+gdn2", lambda: tc.layer_run(tc.TC_CONV5S2, h1[0], h1[1], Pk["conv2"], E.conv2.bias, gdn=Pk["gdn2"], gdn_mode=1, split3=True))
    y = run("conv3", lambda: tc.layer_run(tc.TC_CONV5S2, h2[0], h2[1], Pk["conv3"], None, split3=True, out_f32_nchw=True))
    zh, _ = tc.nchw_to_nhwc(torch.round(y), False)
    d1 = run("deconv1+igdn1", lambda: tc.layer_run(tc.TC_DECONV5S2, zh, None, Pk["deconv1"], D.deconv1.bias, gdn=Pk["igdn1"], gdn_mode=2))
    d2 = run("deconv2+igdn2", lambda: tc.layer_run(tc.TC_DECONV5S2, d1[0], None, Pk["deconv2"], D.deconv2.bias, gdn=Pk["igdn2"], gdn_mode=2))
    r = run("deconv3", lambda: tc.layer_run(tc.TC_DECONV3, d2[0], None, Pk["deconv3"], D.deconv3.bias, out_f32_nchw=True))
    r = run("deconv3+clamp+sse", lambda: tc.layer_run(tc.TC_DECONV3, d2[0], None, Pk["deconv3"], D.deconv3.bias, out_f32_nchw=True, ref_image=x))
