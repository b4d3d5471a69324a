"""GPU parity of each libb2c operator (called through the C ABI via the autograd layer) against the
CPU oracle / plain torch fp32 ops on the same seeded inputs, including ragged shapes."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.fixture(scope="module")
def B2F():
    import iclr_17_compression_b200.functional as f
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return f


CONV_CASES = [
    # B, Cin, H, W, Cout, k, s, p
    (1, 3, 64, 96, 128, 9, 4, 4),
    (2, 128, 32, 48, 128, 5, 2, 2),
    (1, 32, 16, 16, 32, 5, 2, 2),
    (2, 5, 19, 23, 70, 5, 2, 2),       # ragged everything
    (1, 3, 37, 41, 130, 9, 4, 4),
    (3, 7, 9, 9, 3, 3, 1, 1),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d_fwd_bwd(B2F, case):
    B, Cin, H, W, Cout, k, s, p = case
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) * 0.1
    b = torch.randn(Cout, generator=g)
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    yr = F.conv2d(xr, wr, br, stride=s, padding=p)
    go = torch.randn(yr.shape, generator=g)
    yr.backward(go)
    xc, wc, bc = (t.cuda().requires_grad_(True) for t in (x, w, b))
    yc = B2F.conv2d(xc, wc, bc, s, p)
    yc.backward(go.cuda())
    assert yc.shape == yr.shape
    assert _rel(yc, yr) < 2e-6
    assert _rel(xc.grad, xr.grad) < 2e-6
    assert _rel(wc.grad, wr.grad) < 1e-5
    assert _rel(bc.grad, br.grad) < 1e-5
    y2 = B2F.conv2d(xc.detach(), wc.detach(), None, s, p)
    assert _rel(y2, F.conv2d(x, w, None, stride=s, padding=p)) < 2e-6


CONVT_CASES = [
    # B, Cin, H, W, Cout, k, s, p, op
    (2, 128, 8, 12, 128, 5, 2, 2, 1),
    (1, 128, 16, 24, 3, 9, 4, 4, 3),
    (1, 32, 8, 8, 32, 5, 2, 2, 1),
    (2, 6, 7, 5, 70, 5, 2, 2, 1),
    (1, 20, 5, 9, 2, 9, 4, 4, 3),
    (2, 9, 6, 6, 130, 3, 1, 1, 0),
    (1, 4, 6, 7, 5, 5, 2, 2, 0),       # out_pad 0 -> phases of different size
]


@pytest.mark.parametrize("case", CONVT_CASES)
def test_conv_transpose2d_fwd_bwd(B2F, case):
    B, Cin, H, W, Cout, k, s, p, op = case
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cin, Cout, k, k, generator=g) * 0.1
    b = torch.randn(Cout, generator=g)
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    yr = F.conv_transpose2d(xr, wr, br, stride=s, padding=p, output_padding=op)
    go = torch.randn(yr.shape, generator=g)
    yr.backward(go)
    xc, wc, bc = (t.cuda().requires_grad_(True) for t in (x, w, b))
    yc = B2F.conv_transpose2d(xc, wc, bc, s, p, op)
    yc.backward(go.cuda())
    assert yc.shape == yr.shape
    assert _rel(yc, yr) < 2e-6
    assert _rel(xc.grad, xr.grad) < 2e-6
    assert _rel(wc.grad, wr.grad) < 1e-5
    assert _rel(bc.grad, br.grad) < 1e-5


@pytest.mark.parametrize("C,HW,inverse", [(128, (16, 24), False), (128, (16, 24), True), (32, (7, 5), False),
                                          (70, (9, 3), True), (192, (8, 8), False)])
def test_gdn_fwd_bwd(B2F, C, HW, inverse):
    from oracle import balle17_oracle as O
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, C, *HW, generator=g)
    beta = torch.sqrt(0.5 + torch.rand(C, generator=g))
    gamma = torch.sqrt(0.1 * torch.eye(C) + 0.02 * torch.rand(C, C, generator=g) + O.PEDESTAL)
    beta[:3] = O.BETA_BOUND * 0.5                  # exercise LowerBound (GDN.py:18-24)
    gamma.view(-1)[5:40:7] = O.GAMMA_BOUND * 0.25
    xr, br, gr = (t.clone().requires_grad_(True) for t in (x, beta, gamma))
    yr = O.gdn(xr, br, gr, inverse)
    go = torch.randn(yr.shape, generator=g)
    yr.backward(go)
    xc, bc, gc = (t.cuda().requires_grad_(True) for t in (x, beta, gamma))
    yc = B2F.gdn(xc, bc, gc, inverse)
    yc.backward(go.cuda())
    assert _rel(yc, yr) < 2e-6
    assert _rel(xc.grad, xr.grad) < 1e-5
    assert _rel(bc.grad, br.grad) < 1e-4
    assert _rel(gc.grad, gr.grad) < 1e-4
    # LowerBound mask: identical zero pattern
    assert torch.equal(bc.grad.cpu() == 0, br.grad == 0)
    assert torch.equal(gc.grad.cpu() == 0, gr.grad == 0)


@pytest.mark.parametrize("training", [False, True])
@pytest.mark.parametrize("C,HW", [(128, (4, 6)), (32, (5, 3)), (8, (16, 16))])
def test_quant_entropy_fwd_bwd(B2F, training, C, HW):
    from oracle import balle17_oracle as O
    g = torch.Generator().manual_seed(4)
    sd = O.perturb_state_dict(O.init_state_dict(C, seed=5))
    names = ["f1.h", "f1.b", "f1.a", "f2.h", "f2.b", "f2.a", "f3.h", "f3.b", "f3.a", "f4.h", "f4.b"]
    params = [sd["bitEstimator." + n] for n in names]
    y = torch.randn(3, C, *HW, generator=g) * 3
    y.view(-1)[:6] = torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5, 40.0])      # half-to-even ties + a tail value
    noise = (torch.rand(y.shape, generator=g) - 0.5) if training else None
    # oracle
    yr = y.clone().requires_grad_(True)
    pr = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k.startswith("bitEstimator.")}
    zr = O.quantize(yr, training, noise)
    bits_r = O.estimate_bits(zr, pr)
    (bits_r * 0.37).backward()
    # cuda
    yc = y.cuda().requires_grad_(True)
    pc = [p.cuda().requires_grad_(True) for p in params]
    zc, total, image_bits, plane_bits = B2F.quant_entropy(yc, noise.cuda() if training else None, pc)
    (total * 0.37).backward()
    assert torch.equal(zc.cpu(), zr.detach())                                # bit-exact quantizer
    assert abs(total.item() - bits_r.item()) <= 2e-6 * bits_r.item()
    assert abs(image_bits.sum().item() - plane_bits.sum().item()) <= 1e-5 * bits_r.item()
    for n, p in zip(names, pc):
        assert _rel(p.grad, pr["bitEstimator." + n].grad) < 2e-4, n
    if training:
        assert _rel(yc.grad, yr.grad) < 1e-4
    else:
        assert float(yc.grad.abs().max()) == 0.0


def test_entropy_floor_and_clamp(B2F):
    """p underflows in the tails: bits/element = -log2(1e-10) = 33.22 (model.py:73)."""
    from oracle import balle17_oracle as O
    sd = O.perturb_state_dict(O.init_state_dict(8, seed=1))
    names = ["f1.h", "f1.b", "f1.a", "f2.h", "f2.b", "f2.a", "f3.h", "f3.b", "f3.a", "f4.h", "f4.b"]
    params = [sd["bitEstimator." + n].cuda() for n in names]
    y = torch.full((1, 8, 2, 2), 1.0e7, device="cuda")
    _, total, _, _ = B2F.quant_entropy(y, None, params)
    assert abs(total.item() - 32 * 33.219280948) < 1e-2


def test_distortion_fwd_bwd(B2F):
    g = torch.Generator().manual_seed(6)
    x = torch.rand(3, 3, 20, 28, generator=g)
    r = x + 0.6 * torch.randn(x.shape, generator=g)
    rr = r.clone().requires_grad_(True)
    clip_r = rr.clamp(0., 1.)
    sse_r = ((rr - x) ** 2).sum()
    gc = torch.randn(x.shape, generator=g)
    (clip_r * gc).sum().backward(retain_graph=True)
    (sse_r * 0.01).backward()
    rc = r.cuda().requires_grad_(True)
    clipped, sse_total, sse = B2F.distortion(rc, x.cuda())
    ((clipped * gc.cuda()).sum() + sse_total * 0.01).backward()
    assert torch.equal(clipped.cpu(), clip_r.detach())
    assert abs(sse_total.item() - sse_r.item()) < 1e-5 * sse_r.item()
    ref_c = ((clip_r.detach() - x) ** 2).flatten(1).sum(1)
    assert _rel(sse[:, 1], ref_c) < 1e-5
    assert _rel(rc.grad, rr.grad) < 1e-6


def test_errors_are_loud(B2F):
    from iclr_17_compression_b200._lib import B2CError
    with pytest.raises(B2CError):
        B2F.conv2d(torch.rand(1, 3, 8, 8), torch.rand(4, 3, 3, 3), None, 1, 1)        # CPU tensor
    with pytest.raises(B2CError):
        B2F.conv2d(torch.rand(1, 3, 8, 8).cuda(), torch.rand(4, 2, 3, 3).cuda(), None, 1, 1)   # channel mismatch
    with pytest.raises(B2CError):
        B2F.conv2d(torch.rand(1, 3, 8, 8).cuda().half(), torch.rand(4, 3, 3, 3).cuda().half(), None, 1, 1)


@pytest.mark.parametrize("scale", [3.0, 30.0])
def test_entropy_lut_kernel_matches_direct_kernel(scale):
    """table-driven inference entropy pass == the direct kernel: identical z / NHWC bf16 z, per-image bits equal up to
    summation order, including |z| beyond the table (scale 30 -> |z| up to ~120) and the fp32 CPU oracle within 1e-5."""
    from iclr_17_compression_b200 import functional as F
    from oracle import balle17_oracle as O
    sd = O.perturb_state_dict(O.init_state_dict(128, seed=234))
    params = [sd[f"bitEstimator.f{k}.{n}"].reshape(-1).cuda() for k in (1, 2, 3) for n in ("h", "b", "a")] + \
             [sd["bitEstimator.f4.h"].reshape(-1).cuda(), sd["bitEstimator.f4.b"].reshape(-1).cuda()]
    g = torch.Generator().manual_seed(3)
    y = (torch.randn(3, 128, 12, 20, generator=g) * scale).cuda()
    lut = F.entropy_lut(params)
    z0, zh0, b0 = F.quant_entropy_infer(y, params)
    z1, zh1, b1 = F.quant_entropy_infer(y, params, lut=lut)
    assert torch.equal(z0, z1) and torch.equal(zh0, zh1)
    assert torch.allclose(b0, b1, rtol=1e-6)
    ref = torch.stack([O.estimate_bits(torch.round(y[i:i + 1].cpu()), sd) for i in range(3)])
    assert torch.allclose(b1.cpu(), ref, rtol=2e-5)


@pytest.mark.parametrize("C,H,W", [(128, 12, 20), (192, 8, 36), (256, 4, 4), (32, 6, 10), (64, 5, 7)])
def test_sigmoid_binarize_infer_layouts(C, H, W):
    """both kernel versions (C % 64 == 0 and HW % 4 == 0 -> 128-pixel swizzled tile; else 32-pixel padded tile):
    sigmoid within 1 ulp of torch, code = (sigmoid > 0.5) of the kernel's own sigmoid, NHWC bf16 copy == NCHW code"""
    from iclr_17_compression_b200 import functional as F
    g = torch.Generator().manual_seed(C + H)
    y = (torch.randn(3, C, H, W, generator=g) * 2).cuda()
    y.view(-1)[:7] = torch.tensor([0.0, -0.0, 1e-8, -1e-8, 50.0, -50.0, 3e-4])
    code, sig, ch = F.sigmoid_binarize_infer(y)
    assert torch.allclose(sig, torch.sigmoid(y), rtol=3e-7, atol=1e-30)
    assert torch.equal(code, (sig > 0.5).float())
    assert torch.equal(ch, code.permute(0, 2, 3, 1).to(torch.bfloat16))


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_msssim_matches_reference_golden(tag):
    """GPU ssim / ms_ssim (models/ms_ssim_torch.py mirror) vs values produced by the reference's implementation:
    even and odd sizes (the avg_pool padding rule), data_range 1 and 255, per-image and averaged, CPU-tensor inputs."""
    import os
    import numpy as np
    from golden_util import GOLDEN_DIR
    from oracle import balle17_oracle as O
    from iclr_17_compression_b200.models import ms_ssim, ssim, MS_SSIM
    g = np.load(os.path.join(GOLDEN_DIR, "msssim_ref.npz"))
    B, H, W, seed = [int(v) for v in g[f"{tag}/shape"]]
    x, y = O.synth_pair(B, H, W, seed)
    xc, yc = x.cuda(), y.cuda()
    v = ms_ssim(yc, xc, data_range=1.0, size_average=False)
    assert v.is_cuda and tuple(v.shape) == (B,)
    np.testing.assert_allclose(v.cpu().numpy(), g[f"{tag}/ms_ssim"], rtol=2e-5)
    m = ms_ssim(y, x, data_range=1.0, size_average=True)                 # train.py:178 passes CPU tensors
    assert not m.is_cuda and abs(m.item() - float(g[f"{tag}/ms_ssim_mean"])) <= 2e-5 * float(g[f"{tag}/ms_ssim_mean"])
    sv, cs = ssim(yc, xc, data_range=1.0, size_average=False, full=True)
    np.testing.assert_allclose(sv.cpu().numpy(), g[f"{tag}/ssim"], rtol=2e-5)
    np.testing.assert_allclose(cs.cpu().numpy(), g[f"{tag}/cs"], rtol=2e-5)
    v255 = MS_SSIM(data_range=255, size_average=False)(yc * 255, xc * 255)
    np.testing.assert_allclose(v255.cpu().numpy(), g[f"{tag}/ms_ssim_255"], rtol=5e-5)
    with pytest.raises(Exception):
        ms_ssim(yc[:, :, :40, :40], xc[:, :, :40, :40], data_range=1.0)       # 40 -> 20 -> 10 < 11-tap window (the reference fails too)


@pytest.mark.parametrize("world", [1, 2, 4])
def test_allreduce_symm_peer_form_single_device(world):
    """b2c_allreduce_symm, peer form (include/b2c.h): `world` buffers on ONE device stand in for the ranks' symmetric
    copies; each "rank" launches its slice.  Result on every copy == mean in rank order (bit-exact: fixed order)."""
    import ctypes as ct
    from iclr_17_compression_b200 import _lib
    lib = _lib.load()
    n = 4 * world * 1237
    g = torch.Generator(device="cuda").manual_seed(7)
    bufs = [torch.randn(n, device="cuda", generator=g) for _ in range(world)]
    want = torch.zeros(n, device="cuda")
    for b in bufs:
        want = want + b
    want = want * (1.0 / world)
    ptrs = (ct.c_void_p * world)(*[b.data_ptr() for b in bufs])
    st = ct.c_void_p(torch.cuda.current_stream().cuda_stream)
    for r in range(world):
        _lib.check(lib.b2c_allreduce_symm(ct.cast(ptrs, ct.POINTER(ct.c_void_p)), ct.c_void_p(0), r, world, n,
                                          ct.c_float(1.0 / world), st), "allreduce_symm")
    torch.cuda.synchronize()
    for b in bufs:
        assert torch.equal(b, want)
    # loud errors: count not a multiple of 4 * world, rank out of range
    assert lib.b2c_allreduce_symm(ct.cast(ptrs, ct.POINTER(ct.c_void_p)), ct.c_void_p(0), 0, world, n + 1,
                                  ct.c_float(1.0), st) != 0
    assert lib.b2c_allreduce_symm(ct.cast(ptrs, ct.POINTER(ct.c_void_p)), ct.c_void_p(0), world, world, n,
                                  ct.c_float(1.0), st) != 0


def test_bit_estimator_module_forward_matches_reference_golden():
    """model.bitEstimator(x) and Bitparm(x) are module calls in the reference (bitEstimator.py:20-42,
    model.py:71-72); values and gradients against the real reference classes (tests/golden/cdf_W1_n32.npz)."""
    import os
    import numpy as np
    import iclr_17_compression_b200 as P
    from golden_util import GOLDEN_DIR, weights
    g = np.load(os.path.join(GOLDEN_DIR, "cdf_W1_n32.npz"))
    N = int(g["N"])
    be = P.BitEstimator(N)
    be.load_state_dict({k[len("bitEstimator."):]: v for k, v in weights(N, "W1").items() if k.startswith("bitEstimator.")})
    be = be.cuda()
    gen = torch.Generator().manual_seed(4004)
    x = ((torch.rand(2, N, 5, 7, generator=gen) - 0.5) * 16.0).cuda().requires_grad_(True)
    w = torch.randn(2, N, 5, 7, generator=gen).cuda()
    cdf = be(x)
    np.testing.assert_allclose(cdf.detach().cpu().numpy(), g["cdf"], rtol=2e-6, atol=2e-7)
    np.testing.assert_allclose(be.f1(x).detach().cpu().numpy(), g["f1"], rtol=2e-6, atol=2e-7)
    (cdf * w).sum().backward()
    np.testing.assert_allclose(x.grad.cpu().numpy(), g["dx"], rtol=1e-4, atol=1e-7)
    for k, p in be.named_parameters():
        assert _rel(p.grad, torch.from_numpy(g["d/" + k])) < 1e-5, k
    # the reference's own usage pattern (model.py:71-73) agrees with the fused kernel's bit count
    m = P.ImageCompressor(N).cuda().eval()
    m.bitEstimator.load_state_dict(be.state_dict())
    z = torch.round((torch.rand(2, N, 4, 6, generator=gen) - 0.5) * 10.0).cuda()
    prob = m.bitEstimator(z + 0.5) - m.bitEstimator(z - 0.5)
    bits = torch.sum(torch.clamp(-1.0 * torch.log(prob + 1e-10) / 0.6931471805599453, 0, 50))
    _, total, _, _ = m.bitEstimator.bits(z, None)
    assert abs(bits.item() - total.item()) <= 1e-5 * total.item()
