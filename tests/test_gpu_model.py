"""End-to-end GPU parity of ImageCompressor (through the C ABI) against the golden vectors produced by
the real reference and against the CPU oracle; plus size-independent properties at full Kodak size."""
import numpy as np
import pytest
import torch

from golden_util import Case, golden_names

pytestmark = pytest.mark.gpu

# north-star bars
LATENT_BOUNDARY_TOL = 1e-4
REL_BPP_MSE = 1e-3
REL_GRAD = 1e-2


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _model(case, **kw):
    import iclr_17_compression_b200 as P
    m = P.ImageCompressor(case.N, **kw)
    m.load_state_dict(case.sd)
    return m.cuda()


def check_latents(z_gpu, z_ref, y64):
    """bit-exact except where the fp64 pre-round value is within 1e-4 of a .5 boundary."""
    z_gpu, z_ref = z_gpu.cpu(), z_ref.cpu()
    diff = z_gpu != z_ref
    if diff.any():
        frac = (y64.cpu() - torch.floor(y64.cpu()) - 0.5).abs()
        far = diff & (frac > LATENT_BOUNDARY_TOL)
        assert not far.any(), f"{int(far.sum())} latents differ away from a rounding boundary"
    return int(diff.sum())


@pytest.mark.parametrize("name", golden_names("eval_"))
def test_eval_matches_reference_golden(name):
    from oracle import balle17_oracle as O
    c = Case(name)
    m = _model(c).eval()
    with torch.no_grad():
        clipped, z, bpp = m(c.x.cuda())
    g = c.g
    with torch.no_grad():
        o64 = O.forward(c.x.double(), O.to_dtype(c.sd, torch.float64))
    check_latents(z, torch.from_numpy(g["z_i8"].astype(np.float32)), o64["y"])
    assert abs(bpp.item() - float(g["bpp"])) <= REL_BPP_MSE * float(g["bpp"])
    mse = (m.last["sse"][:, 0].sum() / c.x.numel()).item()
    assert abs(mse - float(g["mse"])) <= REL_BPP_MSE * float(g["mse"])
    mse_c = torch.mean((clipped.cpu() - c.x) ** 2).item()
    assert abs(mse_c - float(g["mse_clipped"])) <= REL_BPP_MSE * float(g["mse_clipped"])
    step = int(g["recon_step"])
    rec = m.Decoder(z)
    np.testing.assert_allclose(rec[:, :, ::step, ::step].detach().cpu().numpy(), g["recon_sub"], rtol=1e-3, atol=1e-3)
    assert clipped.dtype == torch.float32 and clipped.is_contiguous() and z.is_contiguous()
    assert tuple(z.shape) == (c.B, c.N, c.H // 16, c.W // 16)


@pytest.mark.parametrize("name", golden_names("train_"))
def test_train_step_matches_reference_golden(name):
    from oracle import balle17_oracle as O
    c = Case(name)
    m = _model(c, return_mode="mse").train()
    m.fast_training = False                                      # the exact-fp32 kernels
    m.noise_override = c.noise.cuda()
    clipped, mse, bpp = m(c.x.cuda())
    loss = c.lam * mse + bpp                                     # train.py:102
    loss.backward()
    g = c.g
    assert abs(loss.item() - float(g["rd_loss"])) <= REL_BPP_MSE * float(g["rd_loss"])
    assert abs(bpp.item() - float(g["bpp"])) <= REL_BPP_MSE * float(g["bpp"])
    _, _, grads = O.rd_loss_and_grads(c.x, c.sd, c.noise, c.lam)
    for k, p in m.named_parameters():
        assert p.grad is not None, k
        r = _rel(p.grad, grads[k])
        assert r <= REL_GRAD, (k, r)
        ref_norm = float(g["g_sum/" + k][2])                    # the real reference's gradient norm
        assert abs(p.grad.double().norm().item() - ref_norm) <= REL_GRAD * ref_norm + 1e-12, k


def test_head_return_tuple_and_eval_with_autograd():
    """model.py:80 tuple order; eval forward with autograd enabled (NewTests/testModelRecon.py:75)."""
    c = Case("eval_W1_U_n32_64")
    m = _model(c).eval()
    clipped, z, bpp = m(c.x.cuda())
    assert clipped.shape == c.x.shape and bpp.dim() == 0
    assert torch.equal(z, torch.round(z))
    bpp.backward()                                               # grads flow to the bit estimator only
    assert m.bitEstimator.f1.h.grad is not None
    assert m.Encoder.conv1.weight.grad is None or float(m.Encoder.conv1.weight.grad.abs().max()) == 0.0


def test_standalone_encoder_decoder_and_training_noise_stream():
    """Encoder/Decoder are public entry points (NewTests/testReconSeperateEandD.py:67-68); training noise
    is drawn with the reference's torch calls from the global CUDA generator (model.py:48-49)."""
    c = Case("train_W1_U_n32_64")
    m = _model(c).train()
    x = c.x.cuda()
    with torch.no_grad():
        y = m.Encoder(x)
        torch.manual_seed(7)
        _, z, _ = m(x)
        torch.manual_seed(7)
        q = torch.zeros(x.size(0), c.N, x.size(2) // 16, x.size(3) // 16).cuda()
        noise = torch.nn.init.uniform_(torch.zeros_like(q), -0.5, 0.5)
        assert torch.equal(z, y + noise)
        r = m.Decoder(torch.round(y))
        assert r.shape == x.shape


def test_bad_shapes_raise():
    c = Case("eval_W1_U_n32_64")
    m = _model(c).eval()
    with pytest.raises(RuntimeError):
        m(torch.rand(1, 3, 40, 64).cuda())


@pytest.mark.parametrize("kind", ["U", "S"])
def test_full_kodak_size_vs_oracle_on_gpu(kind):
    """BASELINE cfg2 size (3x512x768): compare with the oracle's restatement executed by torch on the
    same device in strict fp32 (TF32 off), and check batch-independence (per-image sharding)."""
    from oracle import balle17_oracle as O
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    c = Case("eval_W1_U_n128_64")
    m = _model(c).eval()
    x = O.synth_images(kind, 3, 512, 768, 1002 if kind == "U" else 1003).cuda()
    sd = {k: v.cuda() for k, v in c.sd.items()}
    with torch.no_grad():
        clipped, z, bpp = m(x)
        ref = O.forward(x, sd)
        y64 = O.analysis(x.double(), O.to_dtype(sd, torch.float64))
    ndiff = check_latents(z, ref["z"], y64)
    assert ndiff <= 1e-3 * z.numel()
    assert abs(bpp.item() - ref["bpp"].item()) <= REL_BPP_MSE * ref["bpp"].item()
    mse = (m.last["sse"][:, 0].sum() / x.numel()).item()
    assert abs(mse - ref["mse"].item()) <= REL_BPP_MSE * ref["mse"].item()
    # per-image results do not depend on batching (inference shards per image, no collective)
    with torch.no_grad():
        c1, z1, bpp1 = m(x[1:2])
    assert torch.equal(z1, z[1:2])
    assert torch.equal(c1, clipped[1:2])
    per_image_bpp = m.last["image_bits"][0].item() / (512 * 768)
    assert per_image_bpp > 0


# ------------------------------------------------------------------------------ tensor-core pipeline
@pytest.mark.parametrize("name", [n for n in golden_names("eval_") if "_n128_" in n or "_n192_" in n])
def test_tc_eval_matches_reference_golden(name):
    """eval + no_grad + N=128 runs the tcgen05 pipeline: latents bit-exact away from rounding boundaries
    (bf16x3 analysis), bpp / MSE within 1e-3 relative (bf16 synthesis)."""
    from oracle import balle17_oracle as O
    c = Case(name)
    m = _model(c).eval()
    with torch.no_grad():
        clipped, z, bpp = m(c.x.cuda())
        assert m.last_path.startswith("tcgen05")
        o64 = O.forward(c.x.double(), O.to_dtype(c.sd, torch.float64))
    g = c.g
    check_latents(z, torch.from_numpy(g["z_i8"].astype(np.float32)), o64["y"])
    assert abs(bpp.item() - float(g["bpp"])) <= REL_BPP_MSE * float(g["bpp"])
    mse = (m.last["sse"][:, 0].sum() / c.x.numel()).item()
    assert abs(mse - float(g["mse"])) <= REL_BPP_MSE * float(g["mse"]), (mse, float(g["mse"]))
    mse_c = torch.mean((clipped.cpu() - c.x) ** 2).item()
    assert abs(mse_c - float(g["mse_clipped"])) <= REL_BPP_MSE * float(g["mse_clipped"])


@pytest.mark.parametrize("kind,weights", [("U", "W1"), ("S", "W1"), ("S", "W0")])
def test_tc_full_kodak_size_vs_fp64_oracle(kind, weights):
    """cfg2 size, 3 images: tcgen05 pipeline vs the fp64 oracle on the same device."""
    from oracle import balle17_oracle as O
    from golden_util import weights as W
    import iclr_17_compression_b200 as P
    sd = W(128, weights)
    m = P.ImageCompressor(128)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    x = O.synth_images(kind, 3, 512, 768, 1002 if kind == "U" else 1003).cuda()
    sd64 = {k: v.cuda().double() for k, v in sd.items()}
    with torch.no_grad():
        clipped, z, bpp = m(x)
        assert m.last_path.startswith("tcgen05")
        ref = O.forward(x.double(), sd64)
    ndiff = check_latents(z, ref["z"].float(), ref["y"])
    assert ndiff <= 2e-3 * z.numel()
    assert abs(bpp.item() - ref["bpp"].item()) <= REL_BPP_MSE * ref["bpp"].item()
    mse = (m.last["sse"][:, 0].sum() / x.numel()).item()
    assert abs(mse - ref["mse"].item()) <= REL_BPP_MSE * ref["mse"].item(), (mse, ref["mse"].item())
    # fp32 path on the same input agrees too (path switch is transparent)
    m.fast_inference = False
    with torch.no_grad():
        c2, z2, bpp2 = m(x)
    assert m.last_path == "fp32"
    check_latents(z, z2, ref["y"])
    assert abs(bpp.item() - bpp2.item()) <= REL_BPP_MSE * bpp2.item()
    # batch independence on the tensor-core path
    m.fast_inference = True
    with torch.no_grad():
        c1, z1, _ = m(x[2:3])
    assert torch.equal(z1, z[2:3]) and torch.equal(c1, clipped[2:3])


def test_tc_repacks_after_weight_update():
    c = Case("eval_W1_U_n128_64")
    m = _model(c).eval()
    x = c.x.cuda()
    with torch.no_grad():
        _, z0, bpp0 = m(x)
        m.Encoder.conv3.weight.mul_(1.5)             # in-place update bumps _version -> panels re-packed
        _, z1, bpp1 = m(x)
        m.fast_inference = False
        _, z2, bpp2 = m(x)
    assert not torch.equal(z0, z1)
    assert (z1 != z2).float().mean().item() < 1e-3


def test_host_pipeline_matches_direct_forward():
    """inference.HostPipeline (pinned host in/out, chunked, CUDA graphs) == model(x) on the same images."""
    from iclr_17_compression_b200.inference import HostPipeline
    from oracle import balle17_oracle as O
    c = Case("eval_W1_U_n128_64")
    m = _model(c).eval()
    x = O.synth_images("S", 8, 128, 192, 77)
    with torch.no_grad():
        clipped, z, bpp = m(x.cuda())
        bits = m.last["image_bits"].cpu()
        sse = m.last["sse"].cpu()
    for graphs in (True, False):
        pipe = HostPipeline(m, 8, 128, 192, chunk=2, use_graphs=graphs)
        for _ in range(2):                                   # second call reuses buffers / graphs
            lat, stats = pipe.compress(x.pin_memory())
        assert torch.equal(lat, z.cpu())
        assert torch.allclose(stats[:, 0], bits, rtol=1e-6) and torch.allclose(stats[:, 1:], sse, rtol=1e-5)


def test_host_pipeline_u8_input_matches_fp32_input():
    """compress_u8 (8-bit images over PCIe, ToTensor's / 255 on the device) == compress(u8.float() / 255), bit for bit"""
    from iclr_17_compression_b200.inference import HostPipeline
    c = Case("eval_W1_U_n128_64")
    m = _model(c).eval()
    g = torch.Generator().manual_seed(5)
    u8 = torch.randint(0, 256, (8, 3, 128, 192), generator=g, dtype=torch.uint8)
    x = u8.float().div(255)                                             # torchvision ToTensor
    for graphs in (True, False):
        pipe = HostPipeline(m, 8, 128, 192, chunk=4, use_graphs=graphs)
        lat0, st0 = [t.clone() for t in pipe.compress(x.pin_memory())]
        for _ in range(2):
            lat1, st1 = pipe.compress_u8(u8.pin_memory())
        assert torch.equal(lat0, lat1) and torch.equal(st0, st1)
        lat2, st2 = pipe.compress(x.pin_memory())                      # and the fp32 entry still works afterwards
        assert torch.equal(lat0, lat2)


# ------------------------------------------------------------------------------ bf16 tensor-core training step
@pytest.mark.parametrize("name", [n for n in golden_names("train_") if "_n128_" in n])
def test_tc_train_step_gradients(name):
    """training forward/backward on tensor cores (bf16): loss within 1e-3, every gradient within 1e-2 rel-L2 of
    the oracle (north-star bar), and of the real reference's recorded gradient norms."""
    from oracle import balle17_oracle as O
    c = Case(name)
    m = _model(c, return_mode="mse").train()
    m.fast_training = True
    m.noise_override = c.noise.cuda()
    clipped, mse, bpp = m(c.x.cuda())
    assert m.last_path == "tcgen05-bf16-train"
    loss = c.lam * mse + bpp
    loss.backward()
    g = c.g
    assert abs(loss.item() - float(g["rd_loss"])) <= 2e-3 * float(g["rd_loss"]), (loss.item(), float(g["rd_loss"]))
    assert abs(bpp.item() - float(g["bpp"])) <= REL_BPP_MSE * float(g["bpp"])
    _, _, grads = O.rd_loss_and_grads(c.x, c.sd, c.noise, c.lam)
    worst = {}
    for k, p in m.named_parameters():
        assert p.grad is not None, k
        worst[k] = _rel(p.grad, grads[k])
    bad = {k: v for k, v in worst.items() if v > REL_GRAD}
    assert not bad, bad
    # exact path on the same inputs still available
    m.fast_training = False
    m.zero_grad()
    _, mse2, bpp2 = m(c.x.cuda())
    assert m.last_path == "fp32"
    assert abs(mse2.item() - mse.item()) <= 2e-3 * mse2.item()


@pytest.mark.parametrize("direct", [True, False])
def test_graphed_train_step_matches_eager(direct):
    """GraphedTrainStep (whole-step CUDA graph, flat gradient buffer; autograd-free ``direct`` form and the captured
    autograd form) reproduces the eager tensor-core step."""
    from iclr_17_compression_b200.train_tc import GraphedTrainStep
    c = Case("train_W1_S_n128_64")
    m = _model(c, return_mode="mse").train()
    x, noise = c.x.cuda(), c.noise.cuda()
    m.noise_override = noise
    _, mse, bpp = m(x)
    loss_t = c.lam * mse + bpp
    loss_t.backward()
    eager_loss = loss_t.item()
    eager = {k: p.grad.clone() for k, p in m.named_parameters()}
    del mse, bpp, loss_t, _          # drop the eager autograd graph (its AccumulateGrad nodes live on the legacy stream)
    m.noise_override = None
    step = GraphedTrainStep(m, x, c.lam, direct=direct)
    for _ in range(2):
        loss = step(x, noise)
    torch.cuda.synchronize()
    assert abs(loss.item() - eager_loss) <= 1e-4 * abs(eager_loss)
    for k, p in m.named_parameters():
        assert p.grad.data_ptr() >= step.flat.data_ptr()                      # views of the flat buffer
        assert _rel(p.grad, eager[k]) < 2e-3, (k, _rel(p.grad, eager[k]))      # atomics reorder sums slightly


@pytest.mark.parametrize("N", [192, 256])
def test_tc_wide_latents_vs_fp64_oracle(N):
    """config-5 stand-in: ImageCompressor(192 / 256) on the tensor-core pipeline vs the fp64 oracle (W1-style weights)."""
    from oracle import balle17_oracle as O
    import iclr_17_compression_b200 as P
    sd = O.perturb_state_dict(O.init_state_dict(N, seed=234))
    m = P.ImageCompressor(N)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    x = O.synth_images("S", 2, 256, 384, 1003).cuda()
    sd64 = {k: v.cuda().double() for k, v in sd.items()}
    with torch.no_grad():
        clipped, z, bpp = m(x)
        assert m.last_path.startswith("tcgen05")
        ref = O.forward(x.double(), sd64)
    ndiff = check_latents(z, ref["z"].float(), ref["y"])
    assert ndiff <= 2e-3 * z.numel()
    assert abs(bpp.item() - ref["bpp"].item()) <= REL_BPP_MSE * ref["bpp"].item()
    mse = (m.last["sse"][:, 0].sum() / x.numel()).item()
    assert abs(mse - ref["mse"].item()) <= REL_BPP_MSE * ref["mse"].item(), (mse, ref["mse"].item())
    assert tuple(z.shape) == (2, N, 16, 24)


@pytest.mark.parametrize("shape", [(1, 16, 16), (1, 48, 80), (2, 112, 16), (1, 272, 400), (3, 16, 528)])
def test_tc_ragged_sizes_match_fp32_path(shape):
    """edge cases: images smaller than one 8x16 tile, odd tile counts (a CTA pair gets a fully masked tile),
    tiles that straddle the image border - tensor-core pipeline vs the exact-fp32 kernels on the same input."""
    from oracle import balle17_oracle as O
    B, H, W = shape
    c = Case("eval_W1_U_n128_64")
    m = _model(c).eval()
    x = O.synth_images("U", B, H, W, 4242).cuda()
    with torch.no_grad():
        clipped, z, bpp = m(x)
        assert m.last_path.startswith("tcgen05")
        sse = m.last["sse"].clone()
        m.fast_inference = False
        c2, z2, bpp2 = m(x)
        y64 = O.analysis(x.double(), {k: v.cuda().double() for k, v in c.sd.items()})
    check_latents(z, z2, y64)
    assert abs(bpp.item() - bpp2.item()) <= REL_BPP_MSE * bpp2.item()
    assert _rel(sse[:, 0], m.last["sse"][:, 0]) <= REL_BPP_MSE
    assert _rel(clipped, c2) < 2e-2          # bf16 synthesis vs fp32 (per-pixel error is largest where a latent
    assert (clipped - c2).abs().max().item() < 0.25   # sits on a rounding boundary and legitimately flips)


# ------------------------------------------------------------------------------ fused clamp + Adam (next row N2)
def test_fused_adam_clamp_matches_torch_adam():
    """train.py:106-112: p.grad.data.clamp_(-5, 5) on every tensor, then optim.Adam.step() - one libb2c launch must
    reproduce torch's parameters over several steps (including a learning-rate change, train.py:69-81)."""
    from iclr_17_compression_b200.optim import FusedAdamClamp
    import iclr_17_compression_b200 as P
    torch.manual_seed(5)
    a = P.ImageCompressor(128).cuda()
    b = P.ImageCompressor(128).cuda()
    b.load_state_dict(a.state_dict())
    ref = torch.optim.Adam(a.parameters(), lr=1e-3)
    opt = FusedAdamClamp(b.parameters(), lr=1e-3)
    g = torch.Generator(device="cuda").manual_seed(9)
    for it in range(4):
        if it == 2:
            for o in (ref, opt):
                o.param_groups[0]["lr"] = 3e-4
        for pa, pb in zip(a.parameters(), b.parameters()):
            gr = torch.randn(pa.shape, device="cuda", generator=g) * (20.0 if pa.dim() == 1 else 0.5)   # some beyond +-5
            pa.grad, pb.grad = gr.clone(), gr.clone()
        for pa in a.parameters():
            pa.grad.data.clamp_(-5, 5)
        ref.step()
        opt.step()
    for (k, pa), pb in zip(a.named_parameters(), b.parameters()):
        assert _rel(pb, pa) < 3e-6, (k, _rel(pb, pa))            # fp32 rounding of 4 updates of ~lr each
        assert (pb - pa).abs().max().item() <= 1e-6 * max(1.0, pa.abs().max().item()), k


def test_fused_adam_step_in_ranges_matches_one_launch():
    """begin_step() + step_range() over tensor ranges (GraphedTrainStep's opt-in split optimizer; odd element counts and
    a 3-element tensor shift later tensors off 16-byte alignment: the kernel's scalar path) == one step() launch."""
    from iclr_17_compression_b200.optim import FusedAdamClamp
    torch.manual_seed(11)
    shapes = [(31, 9), (128,), (3,), (128, 128), (1, 128, 1, 1), (7,)]
    pa = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    n = sum(p.numel() for p in pa)
    fa, fb = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    oa, ob = FusedAdamClamp(pa, lr=1e-2, flat_grads=fa), FusedAdamClamp(pb, lr=1e-2, flat_grads=fb)
    g = torch.Generator(device="cuda").manual_seed(2)
    for it in range(3):
        gr = torch.randn(n, device="cuda", generator=g) * 4.0
        fa.copy_(gr), fb.copy_(gr)
        oa.step()
        ob.begin_step()
        ob.step_range(4, 6), ob.step_range(0, 2), ob.step_range(2, 4)
    for a, b in zip(pa, pb):
        assert torch.equal(a, b)
    assert torch.equal(oa.exp_avg, ob.exp_avg) and torch.equal(oa.exp_avg_sq, ob.exp_avg_sq)
    assert int(oa.step_t) == int(ob.step_t) == 3


def test_graphed_train_step_with_fused_optimizer():
    """whole iteration (forward, backward, clamp, Adam) as one CUDA graph replay == graphed step + torch Adam"""
    from iclr_17_compression_b200.train_tc import GraphedTrainStep
    c = Case("train_W1_S_n128_64")
    x, noise = c.x.cuda(), c.noise.cuda()
    ma = _model(c, return_mode="mse").train()
    mb = _model(c, return_mode="mse").train()
    before = {k: p.detach().clone() for k, p in mb.named_parameters()}
    sa = GraphedTrainStep(ma, x, c.lam)
    ref = torch.optim.Adam(ma.parameters(), lr=1e-4)
    sb = GraphedTrainStep(mb, x, c.lam, optimizer="adam")
    for k, p in mb.named_parameters():
        assert torch.equal(p, before[k]), k             # building / warming up the graph must not update anything
    for _ in range(3):
        sa(x, noise)
        for p in ma.parameters():
            p.grad.data.clamp_(-5, 5)
        ref.step()
        sb(x, noise)
    torch.cuda.synchronize()
    # Adam normalises every element's step to ~lr, so the run-to-run noise of the red.add gradient sums flips the
    # sign of near-zero gradients: compare the accumulated updates in relative L2, per tensor
    moved = 0.0
    for (k, pa), pb in zip(ma.named_parameters(), mb.parameters()):
        ua, ub = pa - before[k], pb - before[k]
        assert _rel(ub, ua) < 0.1, (k, _rel(ub, ua))
        assert (pb - pa).abs().max().item() <= 6.5e-4, k              # never more than the 3 steps of 1e-4, both ways
        moved = max(moved, ub.abs().max().item())
    assert moved > 1e-4


# ------------------------------------------------------------------------------ next rows N1: _new and _twoSteps
def _check_code(code_gpu, code_ref, y64):
    """binary code bit-exact except where the fp64 conv3 output is within 1e-4 of the threshold (sigmoid(y) = 0.5 <=> y = 0)"""
    diff = code_gpu.cpu() != code_ref
    far = diff & (y64.cpu().abs() > LATENT_BOUNDARY_TOL)
    assert not far.any(), f"{int(far.sum())} code bits differ away from the threshold"
    return int(diff.sum())


@pytest.mark.parametrize("name", golden_names("new_"))
@pytest.mark.parametrize("fast", [True, False])
def test_new_model_eval_matches_reference_golden(name, fast):
    """ImageCompressor_new (model_new.py:39-63): tensor-core pipeline (N = 128) and exact-fp32 kernels vs the real
    reference's outputs"""
    from golden_util import NewCase
    from oracle import balle17_oracle as O
    import iclr_17_compression_b200 as P
    c = NewCase(name)
    m = P.ImageCompressor_new(c.N)
    m.load_state_dict(c.sd)
    m = m.cuda().eval()
    m.fast_inference = fast
    with torch.no_grad():
        clipped, code, pre = m(c.x.cuda())
        y64 = O.analysis(c.x.double(), O.to_dtype(c.sd, torch.float64))
    assert m.last_path == ("tcgen05" if (fast and c.N in (128, 192, 256)) else "fp32")
    g = c.g
    _check_code(code, torch.from_numpy(g["code_u8"].astype(np.float32)), y64)
    assert (pre.cpu() - torch.from_numpy(g["pre_threshold"])).abs().max().item() < 2e-5
    mse_c = torch.mean((clipped.cpu() - c.x) ** 2).item()
    assert abs(mse_c - float(g["mse_clipped"])) <= (2e-3 if fast else 1e-4) * float(g["mse_clipped"])
    assert set(code.unique().tolist()) <= {0.0, 1.0} and tuple(code.shape) == (c.B, c.N, c.H // 16, c.W // 16)


@pytest.mark.parametrize("name", golden_names("new_"))
def test_new_model_gradients_match_reference_golden(name):
    """training forward/backward of ImageCompressor_new through the straight-through binarizer (exact-fp32 kernels)"""
    from golden_util import NewCase
    from oracle import balle17_oracle as O
    import iclr_17_compression_b200 as P
    c = NewCase(name)
    m = P.ImageCompressor_new(c.N)
    m.load_state_dict(c.sd)
    m = m.cuda().train()
    x = c.x.cuda()
    clipped, z, pre = m(x)
    loss = torch.mean((clipped - x) ** 2) + 0.5 * torch.mean((z - c.target.cuda()) ** 2) + 0.1 * pre.mean()
    loss.backward()
    assert abs(loss.item() - float(c.g["loss"])) <= 1e-4 * float(c.g["loss"])
    _, _, grads = O.new_loss_and_grads(c.x, c.sd, c.target)
    for k, p in m.named_parameters():
        if k.startswith("bitEstimator."):
            assert p.grad is None
            continue
        assert _rel(p.grad, grads[k]) < 1e-3, (k, _rel(p.grad, grads[k]))
        ref_norm = float(c.g["g_sum/" + k][2])
        assert abs(p.grad.double().norm().item() - ref_norm) <= 1e-3 * ref_norm + 1e-12, k


def test_twosteps_model_matches_reference_golden(tmp_path):
    """ImageCompressor_twoSteps: Encoder1 / Decoder1 attribute + state_dict names, same results, step-parsing load_model"""
    from golden_util import NewCase
    from oracle import balle17_oracle as O
    from iclr_17_compression_b200 import model_twoSteps as M2
    c = NewCase("twosteps_W1_U_n128_64")
    m = M2.ImageCompressor_twoSteps(c.N)
    assert sorted(m.state_dict().keys()) == [str(k) for k in c.g["keys"]]
    sd1 = {k.replace("Encoder.", "Encoder1.").replace("Decoder.", "Decoder1."): v for k, v in c.sd.items()}
    f = tmp_path / "iter_1234.pth.tar"
    torch.save(sd1, f)
    assert M2.load_model(m, str(f)) == 1234                           # model_twoSteps.py:30-33
    m = m.cuda().eval()
    with torch.no_grad():
        clipped, z, bpp = m(c.x.cuda())
        y64 = O.analysis(c.x.double(), O.to_dtype(c.sd, torch.float64))
    assert m.last_path.startswith("tcgen05")
    check_latents(z, torch.from_numpy(c.g["z_i8"].astype(np.float32)), y64)
    assert abs(bpp.item() - float(c.g["bpp"])) <= REL_BPP_MSE * float(c.g["bpp"])
    assert torch.equal(m.Encoder1.conv1.weight.cpu(), c.sd["Encoder.conv1.weight"])
    # and one tensor-core training step through the renamed modules
    m.train()
    m.return_mode = "mse"
    _, mse, bpp_t = m(c.x.cuda())
    (8192.0 * mse + bpp_t).backward()
    assert m.last_path == "tcgen05-bf16-train" and m.Encoder1.conv1.weight.grad is not None


# ------------------------------------------------------------------------------ round-1 advisor findings
def test_tc_train_step_latent_mode_cotangents():
    """HEAD return tuple (model.py:80) on the bf16 tensor-core training path: a loss built from the clamped
    reconstruction and the noisy latent must reach every parameter - compared against the exact-fp32 kernels.
    clamp's gradient is discontinuous at 0 and 1 (a bf16-sized change of the reconstruction flips the mask of a pixel
    next to a bound), so the comparison runs with the last layer scaled so that the reconstruction stays inside (0, 1);
    the mask itself is checked with a reconstruction that lies entirely above 1."""
    c = Case("train_W1_S_n128_64")
    x, noise = c.x.cuda(), c.noise.cuda()
    gen = torch.Generator(device="cuda").manual_seed(11)
    wz = torch.randn(noise.shape, device="cuda", generator=gen) * 1e-3

    def model(bias, scale):
        m = _model(c, return_mode="latent").train()
        with torch.no_grad():
            m.Decoder.deconv3.weight.mul_(scale)
            m.Decoder.deconv3.bias.fill_(bias)
        m.noise_override = noise
        return m

    with torch.no_grad():
        probe = model(0.5, 1.0)
        probe.fast_training = False
        r = probe.Decoder(probe.Encoder(x) + noise) - 0.5
        scale = 0.25 / float(r.abs().max())                      # reconstruction within 0.5 +- 0.25
    grads = {}
    for fast in (True, False):
        m = model(0.5, scale)
        m.fast_training = fast
        clipped, z, bpp = m(x)
        assert m.last_path == ("tcgen05-bf16-train" if fast else "fp32")
        assert clipped.requires_grad and z.requires_grad
        assert float(clipped.min()) > 0.0 and float(clipped.max()) < 1.0
        loss = c.lam * torch.mean((clipped - x) ** 2) + bpp + (z * wz).sum()
        loss.backward()
        assert all(p.grad is not None for p in m.parameters())
        grads[fast] = {k: p.grad.clone() for k, p in m.named_parameters()}
    worst = {k: _rel(grads[True][k], grads[False][k]) for k in grads[True]}
    bad = {k: v for k, v in worst.items() if v > REL_GRAD}
    assert not bad, bad
    # distortion through the clamped reconstruction alone: the rate model gets no gradient, the transforms do
    m = model(0.5, scale)
    clipped, z, bpp = m(x)
    torch.mean((clipped - x) ** 2).backward()
    assert float(m.Decoder.deconv3.weight.grad.abs().max()) > 0 and float(m.Encoder.conv1.weight.grad.abs().max()) > 0
    assert float(m.bitEstimator.f1.h.grad.abs().max()) == 0.0
    # reconstruction entirely above 1: clamp passes no gradient at all
    m = model(5.0, scale)
    clipped, z, bpp = m(x)
    assert float(clipped.min()) == 1.0
    torch.mean((clipped - x) ** 2).backward()
    assert float(m.Decoder.deconv3.weight.grad.abs().max()) == 0.0 and float(m.Encoder.conv2.weight.grad.abs().max()) == 0.0


def test_eval_after_fused_optimizer_uses_fresh_weights():
    """eval -> fused clamp + Adam steps (raw-pointer updates: tensor versions do not move) -> eval must re-pack the
    tensor-core weight panels, GDN packs and the entropy table: same outputs as a freshly built model."""
    import iclr_17_compression_b200 as P
    from iclr_17_compression_b200.train_tc import GraphedTrainStep
    from iclr_17_compression_b200.optim import FusedAdamClamp
    c = Case("train_W1_S_n128_64")
    x, noise = c.x.cuda(), c.noise.cuda()
    for in_graph in (True, False):
        m = _model(c, return_mode="latent").eval()
        with torch.no_grad():
            _, z0, bpp0 = m(x)                                   # packs the weights
        m.train()
        if in_graph:
            step = GraphedTrainStep(m, x, c.lam, optimizer="adam")
            step.optimizer.param_groups[0]["lr"] = 1e-2
            for _ in range(3):
                step(x, noise)
        else:
            opt = FusedAdamClamp(m.parameters(), lr=1e-2)
            for _ in range(3):
                m.zero_grad()
                _, _, bpp = m(x)
                bpp.backward()
                opt.step()
        m.eval()
        fresh = P.ImageCompressor(c.N)
        fresh.load_state_dict(m.state_dict())
        fresh = fresh.cuda().eval()
        with torch.no_grad():
            c1, z1, bpp1 = m(x)
            c2, z2, bpp2 = fresh(x)
        assert m.last_path.startswith("tcgen05")
        assert torch.equal(z1, z2) and torch.equal(c1, c2) and bpp1.item() == bpp2.item()
        assert abs(bpp1.item() - bpp0.item()) > 1e-6 * bpp0.item(), "the optimizer steps should have changed the rate"


def test_fused_adam_state_dict_and_missing_grads():
    """optimizer.state_dict() / load_state_dict() carry the moments in torch.optim.Adam's layout (interchangeable with
    it); parameters without a gradient are skipped as torch skips them."""
    from iclr_17_compression_b200.optim import FusedAdamClamp
    torch.manual_seed(3)
    shapes = [(5, 7), (16,), (3, 2, 2)]
    pa = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    pc = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    ref, opt = torch.optim.Adam(pa, lr=1e-2), FusedAdamClamp(pb, lr=1e-2, grad_clip=0.0)
    g = torch.Generator(device="cuda").manual_seed(1)

    def grads(skip=None):
        out = []
        for i, s in enumerate(shapes):
            out.append(None if i == skip else torch.randn(s, device="cuda", generator=g))
        return out

    def apply(params, gs):
        for p, gr in zip(params, gs):
            p.grad = None if gr is None else gr.clone()

    for it in range(3):
        gs = grads()
        apply(pa, gs), apply(pb, gs)
        ref.step(), opt.step()
    sd = opt.state_dict()
    assert set(sd["state"]) == {0, 1, 2} and float(sd["state"][0]["step"]) == 3.0
    for i in range(3):
        assert _rel(sd["state"][i]["exp_avg"], ref.state_dict()["state"][i]["exp_avg"]) < 1e-6
        assert _rel(sd["state"][i]["exp_avg_sq"], ref.state_dict()["state"][i]["exp_avg_sq"]) < 1e-6
    # resume a NEW fused optimizer from torch Adam's own checkpoint, continue both
    for p, q in zip(pc, pa):
        p.data.copy_(q.data)
    opt2 = FusedAdamClamp(pc, lr=5e-2, grad_clip=0.0)
    opt2.load_state_dict(ref.state_dict())
    assert opt2.param_groups[0]["lr"] == 1e-2
    gs = grads()
    apply(pa, gs), apply(pc, gs)
    ref.step(), opt2.step()
    for a, cpar in zip(pa, pc):
        assert _rel(cpar, a) < 1e-6
    # a parameter without a gradient: untouched, moments untouched (torch.optim.Adam skips it)
    before = pb[1].detach().clone()
    m_before = opt.exp_avg.clone()
    apply(pb, grads(skip=1))
    opt.step()
    assert torch.equal(pb[1], before)
    o = shapes[0][0] * shapes[0][1]
    assert torch.equal(opt.exp_avg[o:o + 16], m_before[o:o + 16]) and not torch.equal(opt.exp_avg[:o], m_before[:o])


def test_latent_range_guard_of_the_bf16_synthesis_path():
    """|z| > 256 is not exact in bf16: the tensor-core forward raises its sticky flag, the next call fails loudly and
    the model falls back to the exact-fp32 kernels; decode_tc on caller-supplied latents routes by range."""
    from iclr_17_compression_b200._lib import B2CError
    c = Case("eval_W1_U_n128_64")
    m = _model(c).eval()
    x = c.x.cuda()
    with torch.no_grad():
        m(x)
        torch.cuda.synchronize()
        m.check_latent_range()                                   # normal latents: nothing raised
        m.Encoder.conv3.weight.mul_(400.0)                        # latents far beyond +-256
        _, z, _ = m(x)
        assert m.last_path.startswith("tcgen05") and float(z.abs().max()) > 256
        torch.cuda.synchronize()
        with pytest.raises(B2CError):
            m(x)
        assert m.fast_inference is False
        c_ok, z_ok, _ = m(x)                                     # exact-fp32 kernels now
        assert m.last_path == "fp32" and float((z_ok - z).abs().max()) <= 1.0   # x400 latents: bf16x3 vs fp32 may flip a rounding
        # caller-supplied latents: beyond the range the fp32 synthesis transform runs, below it the tensor-core one
        r_big = m.decode_tc(z)
        assert _rel(r_big, m.Decoder(z)) < 1e-6
        z_small = torch.round(z / 400.0)
        assert _rel(m.decode_tc(z_small), m.Decoder(z_small)) < 2e-2


def test_tc_random_shapes_and_repeated_launches():
    """Protocol stress of the persistent tensor-core kernels (per-warp mbarrier hand-offs, TMEM operand rings, resident
    boxes): random batch / image sizes - odd tile counts give one CTA of a pair a fully masked tile, small images leave
    most CTAs without work - and many back-to-back launches of one shape; every run against the exact-fp32 kernels."""
    import random
    from oracle import balle17_oracle as O
    c = Case("eval_W1_U_n128_64")
    m = _model(c).eval()
    sd64 = {k: v.cuda().double() for k, v in c.sd.items()}
    rng = random.Random(7)
    shapes = [(rng.randint(1, 5), 16 * rng.randint(1, 24), 16 * rng.randint(1, 30)) for _ in range(10)] + [(7, 80, 1040)]
    with torch.no_grad():
        for (B, H, W) in shapes:
            x = O.synth_images("U", B, H, W, 1000 + H + W).cuda()
            m.fast_inference = True
            clipped, z, bpp = m(x)
            assert m.last_path.startswith("tcgen05")
            sse = m.last["sse"].clone()
            m.fast_inference = False
            c2, z2, bpp2 = m(x)
            y64 = O.analysis(x.double(), sd64)
            check_latents(z, z2, y64)
            assert abs(bpp.item() - bpp2.item()) <= REL_BPP_MSE * bpp2.item(), (B, H, W)
            assert _rel(sse[:, 0], m.last["sse"][:, 0]) <= REL_BPP_MSE, (B, H, W)
            assert (clipped - c2).abs().max().item() < 0.25, (B, H, W)
        # the same shape 60 times: every launch must give the same bits
        m.fast_inference = True
        x = O.synth_images("S", 3, 208, 336, 99).cuda()
        ref = None
        for _ in range(60):
            clipped, z, bpp = m(x)
            cur = (z.clone(), clipped.clone(), bpp.clone())
            if ref is None:
                ref = cur
            else:
                assert torch.equal(cur[0], ref[0]) and torch.equal(cur[1], ref[1]) and torch.equal(cur[2], ref[2])


def test_autograd_training_loop_runs_as_graph_replays():
    """The reference's own loop (train.py:97-112: net(input); rd_loss.backward(); clamp; optimizer.step()) on the drop-in
    module: from the second step with the same shape on, forward and backward are CUDA-graph replays under autograd
    (train_tc._StepGraphs).  Same losses / gradients / parameters as the eager tensor-core step, in-place optimizer
    updates are seen by the graphs, a smaller last batch falls back, and a forward that overwrites the saved
    activations before backward fails loudly."""
    from iclr_17_compression_b200._lib import B2CError
    c = Case("train_W1_S_n128_64")
    x, noise = c.x.cuda(), c.noise.cuda()
    models, opts, hist = {}, {}, {}
    for graphed in (True, False):
        m = _model(c, return_mode="mse").train()
        m.graph_training = graphed
        models[graphed], opts[graphed], hist[graphed] = m, torch.optim.Adam(m.parameters(), lr=1e-4), []
    for step in range(5):
        for graphed in (True, False):
            m, opt = models[graphed], opts[graphed]
            m.noise_override = noise
            xs = x if step != 3 else x[:1]                       # step 3: a smaller batch (another shape)
            m.noise_override = noise if step != 3 else noise[:1]
            opt.zero_grad()
            _, mse, bpp = m(xs)
            loss = c.lam * mse + bpp
            loss.backward()
            for p in m.parameters():
                p.grad.data.clamp_(-5, 5)                        # train.py:106-111
            hist[graphed].append((loss.item(), {k: p.grad.clone() for k, p in m.named_parameters()}))
            opt.step()
            want = "tcgen05-bf16-train" + ("-graphed" if graphed and step in (1, 2, 4) else "")
            assert m.last_path == want, (step, graphed, m.last_path)
    for step in range(5):
        (la, ga), (lb, gb) = hist[True][step], hist[False][step]
        assert abs(la - lb) <= 2e-3 * abs(lb), (step, la, lb)
        if step == 0:                                            # later steps: Adam amplifies the atomics' rounding noise
            for k in ga:
                assert _rel(ga[k], gb[k]) < 2e-3, (step, k)
    for (k, pa), pb in zip(models[True].named_parameters(), models[False].parameters()):
        assert (pa - pb).abs().max().item() <= 1.2e-3, k        # 5 Adam steps of 1e-4 each way at most
    # two forwards, then backward of the first: the saved activations are gone -> loud failure
    m = models[True]
    m.noise_override = noise
    _, mse1, bpp1 = m(x)
    _, mse2, bpp2 = m(x)
    with pytest.raises(B2CError):
        (c.lam * mse1 + bpp1).backward()
    (c.lam * mse2 + bpp2).backward()                             # the latest one is fine


def test_small_eval_batches_run_as_one_graph_replay():
    """testKodak feeds one image per call (train.py:162-171): from the second call with the same shape and weights the
    tensor-core eval pipeline is a single CUDA-graph replay with identical results; new weights re-capture; large
    batches and calls under an outer capture stay eager."""
    c = Case("eval_W1_U_n128_64")
    m = _model(c).eval()
    x = c.x.cuda()
    with torch.no_grad():
        a = m(x)
        assert m.last_path == "tcgen05"
        sse_a, bits_a = m.last["sse"].clone(), m.last["image_bits"].clone()
        b = m(x)
        assert m.last_path == "tcgen05-graphed"
        for t, u in zip(a, b):
            assert torch.equal(t, u)
        assert torch.equal(m.last["sse"], sse_a) and torch.equal(m.last["image_bits"], bits_a)
        x2 = torch.rand_like(x)
        c2 = m(x2)                                              # same shape, other data: replay, fresh outputs
        assert m.last_path == "tcgen05-graphed" and not torch.equal(c2[1], b[1])
        assert torch.equal(b[1], a[1])                           # earlier results were copies, not views of graph buffers
        m.graph_inference = False
        d = m(x2)
        assert m.last_path == "tcgen05" and all(torch.equal(t, u) for t, u in zip(c2, d))
        m.graph_inference = True
        m.Encoder.conv3.weight.mul_(1.01)                        # new weight version -> new key -> eager, then re-captured
        e1 = m(x)
        assert m.last_path == "tcgen05"
        e2 = m(x)
        assert m.last_path == "tcgen05-graphed" and torch.equal(e1[1], e2[1]) and not torch.equal(e1[1], a[1])
        big = torch.rand(6, 3, 512, 768, device="cuda")          # 2.4 Mpix: device bound, stays eager
        m(big); m(big); m(big)
        assert m.last_path == "tcgen05"
