"""Warm per-launch time of every kernel on the training step's critical path (batch 8 of 3x256x256): each launch
is captured 20x back to back in a CUDA graph (no host launch cost) and timed with CUDA events; optional clock64
timeline of CTA 0 (argv[1] == 'trace')."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import iclr_17_compression_b200 as P
from iclr_17_compression_b200 import _lib, tc as T, train_tc as TT

lib = _lib.load()
lib.b2c_tc_debug_set_trace.argtypes = [ctypes.c_void_p]
torch.manual_seed(234)
m = P.ImageCompressor(128).cuda().train()
prm = {k: p.detach() for k, p in m.named_parameters()}
B = 8
x = torch.rand(B, 3, 256, 256, device="cuda")
noise = torch.rand(B, 128, 16, 16, device="cuda") - 0.5
st = TT.step_forward(x, noise, prm)
pk = st["pk"]
xs, h1, n1, h2, n2, z, zh, i1, m1, i2, m2, recon = st["saved"]
w1p = T.pack_weights(T.TC_CONV1, prm["Encoder.conv1.weight"], False)
gp1 = T.pack_gdn(prm["Encoder.gdn1.beta"], prm["Encoder.gdn1.gamma"], False)
g_recon = torch.randn_like(recon) * 1e-3
gp_d2, dn2 = T.layer_run(T.TC_CONV1, g_recon, None, pk["d3_dg"], None, gdn=pk["gi2_t"], gdn_mode=4, second_plane=True, saved=(i2, m2))
gp_d1, dn1 = T.layer_run(T.TC_CONV5S2, gp_d2, None, pk["d2_dg"], None, gdn=pk["gi1_t"], gdn_mode=4, second_plane=True, saved=(i1, m1))
g_z = T.layer_run(T.TC_CONV5S2, gp_d1, None, pk["d1_dg"], None, out_f32_nchw=True)
gyh, _ = T.nchw_to_nhwc(g_z, False)
gp_c2, dng2 = T.layer_run(T.TC_DECONV5S2, gyh, None, pk["w3_dg"], None, gdn=pk["gp2_t"], gdn_mode=3, second_plane=True, saved=(h2, n2))
gp_c1, dng1 = T.layer_run(T.TC_DECONV5S2, gp_c2, None, pk["w2_dg"], None, gdn=pk["gp1_t"], gdn_mode=3, second_plane=True, saved=(h1, n1))
g_sse = torch.full((1,), 1e-3, device="cuda")
ws_n = lib.b2c_clamp_sse_workspace_bytes(B, x[0].numel())
ws = torch.empty((ws_n,), device="cuda", dtype=torch.uint8)
clipped, sse = torch.empty_like(recon), torch.empty((B, 2), device="cuda")
bitp = [prm[k].reshape(-1) for k in TT.PARAM_ORDER[19:]]
pp, keep = _lib.ptr_array(bitp)
y = torch.randn(B, 128, 16, 16, device="cuda")
z2, pb, ib = torch.empty_like(y), torch.empty((B, 128), device="cuda"), torch.empty((B,), device="cuda")
s_ = ctypes.c_void_p
def cs(): return s_(torch.cuda.current_stream().cuda_stream)
def p_(t): return s_(t.data_ptr())
zero9 = torch.zeros(128, 3, 9, 9, device="cuda")
tmp5 = torch.zeros(5, 5, 128, 128, device="cuda")
tmp1 = torch.zeros(1, 1, 128, 128, device="cuda")
col = torch.zeros(128, device="cuda")

cases = [
    ("pack_w1", lambda: T.pack_weights(T.TC_CONV1, prm["Encoder.conv1.weight"], False)),
    ("pack_gdn1", lambda: T.pack_gdn(prm["Encoder.gdn1.beta"], prm["Encoder.gdn1.gamma"], False)),
    ("conv1+gdn1 fwd", lambda: T.layer_run(T.TC_CONV1, x, None, w1p, prm["Encoder.conv1.bias"], gdn=gp1, gdn_mode=1, second_plane=True)),
    ("conv2+gdn2 fwd", lambda: T.layer_run(T.TC_CONV5S2, h1, None, pk["w2"], prm["Encoder.conv2.bias"], gdn=pk["gp2"], gdn_mode=1, second_plane=True)),
    ("conv3 fwd", lambda: T.layer_run(T.TC_CONV5S2, h2, None, pk["w3"], None, out_f32_nchw=True)),
    ("entropy fwd", lambda: lib.b2c_quant_entropy_fwd(p_(y), p_(noise), pp, p_(z2), p_(pb), p_(ib), B, 128, 256, cs())),
    ("nchw_to_nhwc z", lambda: T.nchw_to_nhwc(z, False)),
    ("deconv1+igdn1 fwd", lambda: T.layer_run(T.TC_DECONV5S2, zh, None, pk["d1"], prm["Decoder.deconv1.bias"], gdn=pk["gi1"], gdn_mode=2, second_plane=True)),
    ("deconv2+igdn2 fwd", lambda: T.layer_run(T.TC_DECONV5S2, i1, None, pk["d2"], prm["Decoder.deconv2.bias"], gdn=pk["gi2"], gdn_mode=2, second_plane=True)),
    ("deconv3 fwd", lambda: T.layer_run(T.TC_DECONV3, i2, None, pk["d3"], prm["Decoder.deconv3.bias"], out_f32_nchw=True)),
    ("clamp_sse fwd", lambda: lib.b2c_clamp_sse_fwd(p_(recon), p_(x), p_(clipped), p_(sse), B, x[0].numel(), p_(ws), ws_n, cs())),
    ("clamp_sse bwd", lambda: lib.b2c_clamp_sse_bwd(p_(recon), p_(x), s_(0), p_(g_sse), p_(g_recon), recon.numel(), cs())),
    ("dgrad deconv3 (conv1 kind)+igdn2 bwd", lambda: T.layer_run(T.TC_CONV1, g_recon, None, pk["d3_dg"], None, gdn=pk["gi2_t"], gdn_mode=4, second_plane=True, saved=(i2, m2))),
    ("dgrad deconv2 (conv5s2)+igdn1 bwd", lambda: T.layer_run(T.TC_CONV5S2, gp_d2, None, pk["d2_dg"], None, gdn=pk["gi1_t"], gdn_mode=4, second_plane=True, saved=(i1, m1))),
    ("dgrad deconv1 (conv5s2)", lambda: T.layer_run(T.TC_CONV5S2, gp_d1, None, pk["d1_dg"], None, out_f32_nchw=True)),
    ("add g_z+dz", lambda: g_z + g_z),
    ("nchw_to_nhwc g_y", lambda: T.nchw_to_nhwc(g_z, False)),
    ("dgrad conv3 (deconv5s2)+gdn2 bwd", lambda: T.layer_run(T.TC_DECONV5S2, gyh, None, pk["w3_dg"], None, gdn=pk["gp2_t"], gdn_mode=3, second_plane=True, saved=(h2, n2))),
    ("dgrad conv2 (deconv5s2)+gdn1 bwd", lambda: T.layer_run(T.TC_DECONV5S2, gp_c2, None, pk["w2_dg"], None, gdn=pk["gp1_t"], gdn_mode=3, second_plane=True, saved=(h1, n1))),
    ("wgrad9 conv1", lambda: T.wgrad9(gp_c1, x, out=zero9)),
    ("wgrad 5x5 conv2 (8x32x32)", lambda: T.wgrad(gp_c2, h1, 5, 2, 2, tmp=tmp5)),
    ("wgrad 5x5 conv3 (8x16x16)", lambda: T.wgrad(gyh, h2, 5, 2, 2, tmp=tmp5)),
    ("wgrad 1tap (8x64x64)", lambda: T.wgrad(dng1, h1, 1, 1, 0, tmp=tmp1)),
    ("colsum 8x64x64", lambda: TT._colsum(gp_c1, col)),
    ("xsq 8x64x64", lambda: TT._xsq(h1, n1, False)),
]
R = 20
tot = 0.0
for name, fn in cases:
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(R):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / R * 1e3
    print(f"{us:8.2f} us  {name}")
    if len(sys.argv) > 1 and sys.argv[1] == "trace" and "fwd" in name or "dgrad" in name:
        buf = torch.zeros(3 * 64, dtype=torch.int64, device="cuda")
        lib.b2c_tc_debug_set_trace(ctypes.c_void_p(buf.data_ptr()))
        fn(); torch.cuda.synchronize()
        lib.b2c_tc_debug_set_trace(ctypes.c_void_p(0))
        t = buf.cpu().view(3, 64)
        nz = [int(v) for v in t.flatten() if v > 0]
        if nz:
            t0 = min(nz)
            for role, nm in enumerate(["producer", "mma", "worker"]):
                vals = [int(v) - t0 for v in t[role] if v > 0]
                print(f"      {nm:8s}", vals)
