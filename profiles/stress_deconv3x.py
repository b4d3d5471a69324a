"""Random-shape stress of the x-folded last synthesis layer (all three widths; heights / widths from 1): values, clamp,
squared errors and distortion gradient against conv_transpose2d of the bf16-rounded operands."""
import sys, random
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from iclr_17_compression_b200 import tc
random.seed(3)
g = torch.Generator().manual_seed(3)
bad = 0
for it in range(40):
    C = random.choice([128, 128, 192, 256])
    B = random.randint(1, 3); H = random.randint(1, 40); W = random.randint(1, 60)
    x = torch.randn(B, H, W, C, generator=g).cuda().to(torch.bfloat16)
    w = (torch.randn(C, 3, 9, 9, generator=g) * 0.05).cuda()
    b = torch.randn(3, generator=g).cuda()
    img = torch.rand(B, 3, 4 * H, 4 * W, generator=g).cuda()
    ref = F.conv_transpose2d(x.float().permute(0, 3, 1, 2).double(), w.to(torch.bfloat16).double(), b.double(), stride=4, padding=4, output_padding=3)
    wp = tc.pack_weights(tc.TC_DECONV3X, w, False)
    y = tc.layer_run(tc.TC_DECONV3X, x, None, wp, b, out_f32_nchw=True)
    cl, sse, gr = tc.layer_run(tc.TC_DECONV3X, x, None, wp, b, out_f32_nchw=True, ref_image=img, grad_scale=0.5)
    e1 = ((y.double() - ref).norm() / ref.norm()).item()
    e2 = (cl.double() - ref.clamp(0, 1)).abs().max().item()
    sref = ((ref - img.double()) ** 2).flatten(1).sum(1)
    e3 = ((sse[:, 0].double() - sref).abs() / sref).max().item()
    e4 = ((gr.double() - 0.5 * (ref - img.double())).norm() / (ref - img.double()).norm()).item()
    ok = e1 < 2e-5 and e2 < 3e-5 and e3 < 1e-4 and e4 < 2e-5
    bad += not ok
    if not ok or it % 10 == 0:
        print(it, C, B, H, W, f"{e1:.1e} {e2:.1e} {e3:.1e} {e4:.1e}", "OK" if ok else "BAD", flush=True)
print("bad", bad)
