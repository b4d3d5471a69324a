#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REAL reference (/root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):

    CUDA_VISIBLE_DEVICES="" python tests/golden/make_golden.py

Shims: (1) a stub ``imageio`` module because
models/basics.py:15 imports it at module scope; (2) ``torch.Tensor.cuda`` is a
no-op because model.py:48 calls ``.cuda()`` unconditionally.  Nothing else of
the reference is touched: modules, init and forward are the reference's own.

Each case stores: the weight set name + per-tensor checksums (so the tests know
their regenerated weights are the reference's), the input seeds, and the
reference outputs (latents, bpp, mse, reconstruction, and for training cases the
rd_loss and a digest of each of the 30 gradient tensors).
"""
import os
import sys
import types

os.environ.setdefault("CUDA_VISIBLE_DEVICES", "")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"

import numpy as np
import torch

sys.modules.setdefault("imageio", types.ModuleType("imageio"))
torch.Tensor.cuda = lambda self, *a, **k: self
sys.path.insert(0, REF)
import model as refmodel  # noqa: E402  (the reference's model.py)

sys.path.insert(0, ROOT)
from oracle import balle17_oracle as O  # noqa: E402  (only for synth inputs + the W1 perturbation)

torch.set_num_threads(os.cpu_count())
torch.backends.mkldnn.enabled = True


def tensor_digest(t: torch.Tensor):
    t = t.detach().double().flatten()
    idx = torch.linspace(0, t.numel() - 1, steps=min(64, t.numel())).long()
    return np.array([t.sum().item(), t.abs().sum().item(), t.norm().item()]), t[idx].numpy()


def build_ref(N, weights):
    torch.manual_seed(234)                       # train.py:39,202
    net = refmodel.ImageCompressor(out_channel_N=N)
    if weights == "W1":
        sd = O.perturb_state_dict({k: v.detach().clone() for k, v in net.state_dict().items()})
        net.load_state_dict(sd)
    return net


def run_case(name, N, weights, kind, B, H, W, xseed, training, nseed=None, lam=8192.0):
    net = build_ref(N, weights)
    x = O.synth_images(kind, B, H, W, xseed)
    out = {"N": N, "weights": weights, "kind": kind, "B": B, "H": H, "W": W, "xseed": xseed,
           "training": int(training), "nseed": -1 if nseed is None else nseed, "lam": lam}
    for k, v in net.state_dict().items():
        s, smp = tensor_digest(v)
        out["w_sum/" + k] = s
        out["w_smp/" + k] = smp
    if not training:
        net.eval()
        with torch.no_grad():
            clipped, z, bpp = net(x)                                  # model.py:47-80, HEAD tuple
            y = net.Encoder(x)
            recon = net.Decoder(z)
        assert torch.equal(torch.round(y), z)
        mse = torch.mean((recon - x).pow(2))                          # model.py:61
        out["z_i8"] = z.numpy().astype(np.int8)
        assert np.array_equal(out["z_i8"].astype(np.float32), z.numpy())
        out["y"] = y.numpy()
    else:
        net.train()
        noise = O.synth_noise(B, N, H, W, nseed)
        # the reference draws its own noise inside forward (model.py:48-49); to inject ours we
        # run the same sub-module composition forward() does, and assert forward() agrees with
        # that composition when given the same generator state.
        y = net.Encoder(x)
        z = y + noise
        recon = net.Decoder(z)
        clipped = recon.clamp(0., 1.)
        mse = torch.mean((recon - x).pow(2))
        prob = net.bitEstimator(z + 0.5) - net.bitEstimator(z - 0.5)
        bits = torch.sum(torch.clamp(-1.0 * torch.log(prob + 1e-10) / np.log(2.0), 0, 50))
        bpp = bits / (B * H * W)
        torch.manual_seed(99)
        c2, z2, bpp2 = net(x)
        torch.manual_seed(99)
        n2 = torch.nn.init.uniform_(torch.zeros(B, N, H // 16, W // 16), -0.5, 0.5)
        assert torch.equal(z2, y + n2), "forward() composition drifted from model.py"
        loss = lam * mse + bpp                                        # train.py:102
        net.zero_grad()
        loss.backward()
        out["rd_loss"] = np.float64(loss.item())
        for k, p in net.named_parameters():
            s, smp = tensor_digest(p.grad)
            out["g_sum/" + k] = s
            out["g_smp/" + k] = smp
        out["z"] = z.detach().numpy()
        out["y"] = y.detach().numpy()
    out["bpp"] = np.float64(bpp.item())
    out["mse"] = np.float64(mse.item())
    out["mse_clipped"] = np.float64(torch.mean((clipped - x).pow(2)).item())   # train.py:172
    r = recon.detach()
    out["recon_sum"] = tensor_digest(r)[0]
    step = max(1, H // 64)
    out["recon_sub"] = r[:, :, ::step, ::step].numpy()
    out["recon_step"] = step
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: bpp={out['bpp']:.6f} mse={out['mse']:.6f} -> {os.path.getsize(path)/1024:.0f} KB")


def run_case_new(name, N, kind, B, H, W, xseed, tseed):
    """ImageCompressor_new (model_new.py): eval outputs and the gradients of a three-output loss (W1 weights)."""
    import model_new as refnew
    torch.manual_seed(234)
    net = refnew.ImageCompressor_new(out_channel_N=N)
    sd = O.perturb_state_dict({k: v.detach().clone() for k, v in net.state_dict().items()})
    net.load_state_dict(sd)
    x = O.synth_images(kind, B, H, W, xseed)
    out = {"N": N, "kind": kind, "B": B, "H": H, "W": W, "xseed": xseed, "tseed": tseed}
    for k, v in net.state_dict().items():
        s_, smp = tensor_digest(v)
        out["w_sum/" + k] = s_
        out["w_smp/" + k] = smp
    net.eval()
    with torch.no_grad():
        clipped, z, pre = net(x)
    out["code_u8"] = z.numpy().astype(np.uint8)
    out["pre_threshold"] = pre.numpy()
    out["clipped_sum"] = tensor_digest(clipped)[0]
    step = max(1, H // 64)
    out["clipped_sub"] = clipped[:, :, ::step, ::step].numpy()
    out["clipped_step"] = step
    out["mse_clipped"] = np.float64(torch.mean((clipped - x).pow(2)).item())
    net.train()
    g = torch.Generator().manual_seed(tseed)
    target = (torch.rand(B, N, H // 16, W // 16, generator=g) > 0.5).float()
    clipped, z, pre = net(x)
    loss = torch.mean((clipped - x) ** 2) + 0.5 * torch.mean((z - target) ** 2) + 0.1 * pre.mean()
    net.zero_grad()
    loss.backward()
    out["loss"] = np.float64(loss.item())
    for k, p in net.named_parameters():
        if p.grad is None:
            continue
        s_, smp = tensor_digest(p.grad)
        out["g_sum/" + k] = s_
        out["g_smp/" + k] = smp
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: ones={z.mean().item():.4f} loss={out['loss']:.6f} -> {os.path.getsize(path)/1024:.0f} KB")


def run_case_twosteps(name, N, kind, B, H, W, xseed):
    """ImageCompressor_twoSteps (model_twoSteps.py): eval outputs with W1 weights under the Encoder1/Decoder1 keys."""
    import model_twoSteps as ref2
    torch.manual_seed(234)
    net = ref2.ImageCompressor_twoSteps(out_channel_N=N)
    sd = O.perturb_state_dict({k.replace("Encoder1.", "Encoder.").replace("Decoder1.", "Decoder."): v.detach().clone()
                               for k, v in net.state_dict().items()})
    net.load_state_dict({k.replace("Encoder.", "Encoder1.").replace("Decoder.", "Decoder1."): v for k, v in sd.items()})
    x = O.synth_images(kind, B, H, W, xseed)
    net.eval()
    with torch.no_grad():
        clipped, z, bpp = net(x)
        y = net.Encoder1(x)
    out = {"N": N, "kind": kind, "B": B, "H": H, "W": W, "xseed": xseed, "keys": np.array(sorted(net.state_dict().keys())),
           "z_i8": z.numpy().astype(np.int8), "y": y.numpy(), "bpp": np.float64(bpp.item()),
           "mse_clipped": np.float64(torch.mean((clipped - x).pow(2)).item())}
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: bpp={out['bpp']:.6f} -> {os.path.getsize(path)/1024:.0f} KB")


if __name__ == "__main__":
    if "--msssim" in sys.argv:          # SURVEY 8f N3 fixture: the reference's own models/ms_ssim_torch.py on seeded pairs
        from models.ms_ssim_torch import ms_ssim as ref_ms_ssim, ssim as ref_ssim
        out = {}
        for tag, (B, H, W, seed) in {"a": (2, 256, 384, 7001), "b": (1, 250, 391, 7002), "c": (3, 176, 176, 7003)}.items():
            x, y = O.synth_pair(B, H, W, seed)
            out[f"{tag}/shape"] = np.array([B, H, W, seed])
            # (size_average=False with B > 1 raises inside the reference - its NaN check `if not v == v` is ambiguous
            #  for a vector - so per-image values are taken one image at a time)
            out[f"{tag}/ms_ssim"] = np.array([ref_ms_ssim(y[i:i + 1], x[i:i + 1], data_range=1.0, size_average=False).item()
                                              for i in range(B)], dtype=np.float32)
            out[f"{tag}/ms_ssim_mean"] = np.float64(ref_ms_ssim(y, x, data_range=1.0, size_average=True).item())   # train.py:178
            sv, cs = ref_ssim(y, x, data_range=1.0, size_average=False, full=True)
            out[f"{tag}/ssim"], out[f"{tag}/cs"] = sv.numpy(), cs.numpy()
            out[f"{tag}/ms_ssim_255"] = np.array([ref_ms_ssim(y[i:i + 1] * 255, x[i:i + 1] * 255, data_range=255,
                                                              size_average=False).item() for i in range(B)], dtype=np.float32)
        np.savez_compressed(os.path.join(HERE, "msssim_ref.npz"), **out)
        print({k: v for k, v in out.items() if "ms_ssim" in k and "mean" in k})
        sys.exit(0)
    if "--next-rows" in sys.argv:      # SURVEY 8f N1 fixtures only (the base fixtures above stay untouched)
        run_case_new("new_W1_S_n128_64", 128, "S", 2, 64, 96, 1003, 4004)
        run_case_new("new_W1_U_n32_64", 32, "U", 1, 64, 64, 1002, 4004)
        run_case_twosteps("twosteps_W1_U_n128_64", 128, "U", 2, 64, 96, 1002)
        sys.exit(0)
    # (name, N, weights, kind, B, H, W, xseed, training, nseed)
    run_case("eval_W0_U_n128_64", 128, "W0", "U", 2, 64, 96, 1002, False)
    run_case("eval_W1_U_n128_64", 128, "W1", "U", 2, 64, 96, 1002, False)
    run_case("eval_W1_S_n128_128", 128, "W1", "S", 1, 128, 192, 1003, False)
    run_case("eval_W1_U_n32_64", 32, "W1", "U", 1, 64, 64, 1002, False)
    run_case("eval_W1_S_n192_64", 192, "W1", "S", 1, 64, 128, 1003, False)
    run_case("train_W0_U_n128_256", 128, "W0", "U", 1, 256, 256, 1001, True, 2002)   # cfg1
    run_case("train_W1_S_n128_64", 128, "W1", "S", 2, 64, 96, 1004, True, 2002)
    run_case("train_W1_U_n32_64", 32, "W1", "U", 2, 64, 64, 1004, True, 2002)
