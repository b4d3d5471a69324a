"""One invocation of the N4 (bitstream) and N5 (input transform) kernels at bandwidth sizes, for ncu:
python profiles/run_next_rows_once.py ; ncu --set full -k regex:'rans|resized' ... python profiles/run_next_rows_once.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch

import iclr_17_compression_b200 as P
from iclr_17_compression_b200 import bitstream as BS
from iclr_17_compression_b200 import datasets as D
from golden_util import weights as W
from oracle import balle17_oracle as O
from oracle import resample_oracle as R

m = P.ImageCompressor(128)
m.load_state_dict(W(128, "W1"))
m = m.cuda().eval()
tables = BS.RansTables.from_bit_estimator(m.bitEstimator)
x = O.synth_images("S", 24, 512, 768, 1003).cuda()
with torch.no_grad():
    _, z, _ = m(x)
out, off, bb, st = BS.encode_latents_raw(z, tables, 8192)
z2, st2 = BS.decode_latents_raw(out, off, tables, *z.shape, 8192)
assert torch.equal(z, z2)

base = [R.synth_photo(1356, 2040, s) for s in range(4)]
imgs = [base[n % 4] for n in range(64)]
torch.manual_seed(5)
params = [D.sample_params(1356, 2040) for _ in range(64)]
buf, desc, mh, mw = D.pack_images(imgs, params, crop_only=False)
y = D.resized_crop_flip_packed(buf.cuda(), desc.cuda(), 256, mh, mw)
torch.cuda.synchronize()
print("ok", float(y.mean()), int(off[-1]))
