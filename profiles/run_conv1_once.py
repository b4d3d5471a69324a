"""A few launches of the CONV1+GDN1 layer alone (24 Kodak-shape images), for `ncu -k regex:conv1 -s 3 -c 1`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import iclr_17_compression_b200 as P
from iclr_17_compression_b200 import tc

torch.manual_seed(234)
m = P.ImageCompressor(128).cuda().eval()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
x = torch.rand(B, 3, 512, 768, device="cuda")
Pk = m._packed_weights()
with torch.no_grad():
    for _ in range(5):
        tc.layer_run(tc.TC_CONV1, x, None, Pk["conv1"], m.Encoder.conv1.bias, gdn=Pk["gdn1"], gdn_mode=1, split3=True)
torch.cuda.synchronize()
print("ok")
