"""The reference's OWN training loop on the drop-in module (train.py:92-112: net(input); rd_loss = lambda * mse + bpp;
optimizer.zero_grad(); rd_loss.backward(); element-wise gradient clamp; optimizer.step()), batch 8 of 3x256x256:
img/s with the autograd path as CUDA-graph replays (B2C_TRAIN_AUTOGRAPH=1, default) and eager (=0)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import iclr_17_compression_b200 as P

torch.manual_seed(234)
net = P.ImageCompressor(128, return_mode="mse").cuda().train()
opt = torch.optim.Adam(net.parameters(), lr=1e-4)
x = torch.rand(8, 3, 256, 256, device="cuda")

def step():
    clipped, mse, bpp = net(x)
    rd_loss = 8192.0 * mse + bpp
    opt.zero_grad()
    rd_loss.backward()
    for group in opt.param_groups:                       # train.py:106-111 clip_gradient(optimizer, 5)
        for p in group["params"]:
            if p.grad is not None:
                p.grad.data.clamp_(-5, 5)
    opt.step()

for _ in range(5):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 50
for _ in range(n):
    step()
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / n * 1e3
print(f"B2C_TRAIN_AUTOGRAPH={os.environ.get('B2C_TRAIN_AUTOGRAPH', '1')} path={net.last_path}: {ms:.3f} ms / step, {8 / ms * 1e3:.0f} img/s")
