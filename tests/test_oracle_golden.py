"""Pins oracle/balle17_oracle.py against outputs of the real reference (tests/golden/*.npz,
produced by tests/golden/make_golden.py from /root/reference).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import balle17_oracle as O
from golden_util import Case, digest, golden_names


@pytest.mark.parametrize("name", golden_names("eval_"))
def test_eval_matches_reference(name):
    c = Case(name)
    c.check_weights()
    with torch.no_grad():
        out = O.forward(c.x, c.sd, training=False)
    g = c.g
    # integer latents: bit-exact
    np.testing.assert_array_equal(out["z"].numpy(), g["z_i8"].astype(np.float32))
    np.testing.assert_allclose(out["y"].numpy(), g["y"], rtol=0, atol=1e-5)
    assert abs(out["bpp"].item() - float(g["bpp"])) <= 1e-6 * abs(float(g["bpp"]))
    assert abs(out["mse"].item() - float(g["mse"])) <= 1e-5 * abs(float(g["mse"]))
    mse_c = torch.mean((out["clipped"] - c.x) ** 2).item()
    assert abs(mse_c - float(g["mse_clipped"])) <= 1e-5 * float(g["mse_clipped"])
    step = int(g["recon_step"])
    np.testing.assert_allclose(out["recon"][:, :, ::step, ::step].numpy(), g["recon_sub"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", golden_names("train_"))
def test_train_and_grads_match_reference(name):
    c = Case(name)
    c.check_weights()
    out, loss, grads = O.rd_loss_and_grads(c.x, c.sd, c.noise, c.lam)
    g = c.g
    np.testing.assert_allclose(out["z"].detach().numpy(), g["z"], rtol=0, atol=1e-5)
    assert abs(loss.item() - float(g["rd_loss"])) <= 1e-5 * abs(float(g["rd_loss"]))
    assert abs(out["bpp"].item() - float(g["bpp"])) <= 1e-6 * abs(float(g["bpp"]))
    for k, gr in grads.items():
        s, smp = digest(gr)
        ref_norm = float(g["g_sum/" + k][2])
        assert abs(s[2] - ref_norm) <= 1e-4 * ref_norm + 1e-12, (k, s[2], ref_norm)
        np.testing.assert_allclose(smp, g["g_smp/" + k], rtol=2e-3, atol=1e-4 * ref_norm / np.sqrt(gr.numel()) + 1e-12, err_msg=k)


def test_fp64_oracle_close_to_fp32():
    c = Case("eval_W1_U_n32_64")
    with torch.no_grad():
        o32 = O.forward(c.x, c.sd)
        o64 = O.forward(c.x.double(), O.to_dtype(c.sd, torch.float64))
    assert (o32["y"].double() - o64["y"]).abs().max() < 1e-4
    assert abs(o32["bpp"].item() - o64["bpp"].item()) < 1e-5 * o64["bpp"].item()


def test_lower_bound_gradient_rule():
    # models/GDN.py:18-24
    x = torch.tensor([0.5, 2.0, 0.5, 2.0], requires_grad=True)
    y = O._LowerBound.apply(x, 1.0)
    y.backward(torch.tensor([1.0, 1.0, -1.0, -1.0]))
    assert x.grad.tolist() == [0.0, 1.0, -1.0, -1.0]


def test_entropy_floor_in_the_tails():
    # p underflows -> -log2(1e-10) = 33.22 bits per element (model.py:73)
    sd = O.init_state_dict(8, seed=1)
    sd = O.perturb_state_dict(sd)
    z = torch.full((1, 8, 2, 2), 1.0e7)
    bits = O.estimate_bits(z, sd).item()
    assert abs(bits - 32 * 33.219280948) < 1e-2


# ------------------------------------------------------------------------------ next rows
@pytest.mark.parametrize("name", golden_names("new_"))
def test_new_model_matches_reference(name):
    """oracle.forward_new / new_loss_and_grads vs the real model_new.ImageCompressor_new (binary code bit-exact)"""
    from golden_util import NewCase
    c = NewCase(name)
    g = c.g
    for k, v in c.sd.items():
        s, smp = digest(v)
        np.testing.assert_allclose(s, g["w_sum/" + k], rtol=1e-12, atol=0, err_msg=k)
    with torch.no_grad():
        out = O.forward_new(c.x, c.sd, training=False)
    np.testing.assert_array_equal(out["z"].numpy(), g["code_u8"].astype(np.float32))
    np.testing.assert_allclose(out["pre_threshold"].numpy(), g["pre_threshold"], rtol=0, atol=1e-6)
    step = int(g["clipped_step"])
    np.testing.assert_allclose(out["clipped"][:, :, ::step, ::step].numpy(), g["clipped_sub"], rtol=1e-4, atol=1e-5)
    _, loss, grads = O.new_loss_and_grads(c.x, c.sd, c.target)
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    for k, gr in grads.items():
        s, smp = digest(gr)
        ref_norm = float(g["g_sum/" + k][2])
        assert abs(s[2] - ref_norm) <= 1e-4 * ref_norm + 1e-12, (k, s[2], ref_norm)


def test_twosteps_matches_reference():
    """model_twoSteps.ImageCompressor_twoSteps == ImageCompressor under Encoder1 / Decoder1 keys"""
    from golden_util import NewCase
    c = NewCase("twosteps_W1_U_n128_64")
    with torch.no_grad():
        out = O.forward(c.x, c.sd, training=False)
    np.testing.assert_array_equal(out["z"].numpy(), c.g["z_i8"].astype(np.float32))
    assert abs(out["bpp"].item() - float(c.g["bpp"])) <= 1e-6 * float(c.g["bpp"])
    keys = sorted(k.replace("Encoder.", "Encoder1.").replace("Decoder.", "Decoder1.") for k in c.sd)
    assert keys == [str(k) for k in c.g["keys"]]


def test_msssim_oracle_matches_reference():
    """oracle.ms_ssim_per_image / ssim_per_image vs the reference's models/ms_ssim_torch.py (tests/golden/msssim_ref.npz)"""
    import os
    from golden_util import GOLDEN_DIR
    g = np.load(os.path.join(GOLDEN_DIR, "msssim_ref.npz"))
    for tag in ("a", "b", "c"):
        B, H, W, seed = [int(v) for v in g[f"{tag}/shape"]]
        x, y = O.synth_pair(B, H, W, seed)
        np.testing.assert_allclose(O.ms_ssim_per_image(y, x, 1.0).numpy(), g[f"{tag}/ms_ssim"], rtol=1e-6)
        sv, cs = O.ssim_per_image(y, x, O._gauss_1d(11, 1.5), 1.0)
        np.testing.assert_allclose(sv.numpy(), g[f"{tag}/ssim"], rtol=1e-6)
        np.testing.assert_allclose(cs.numpy(), g[f"{tag}/cs"], rtol=1e-6)


def test_cdf_oracle_matches_reference_module():
    """BitEstimator.forward / Bitparm.forward as module calls (bitEstimator.py:20-42): the oracle's bit_estimator /
    bitparm against values and gradients produced by the real reference classes (tests/golden/make_cdf_golden.py)."""
    import os
    import numpy as np
    import torch
    from golden_util import GOLDEN_DIR, weights
    from oracle import balle17_oracle as O
    g = np.load(os.path.join(GOLDEN_DIR, "cdf_W1_n32.npz"))
    N = int(g["N"])
    sd = {k: v.clone().requires_grad_(True) for k, v in weights(N, "W1").items() if k.startswith("bitEstimator.")}
    gen = torch.Generator().manual_seed(4004)
    x = ((torch.rand(2, N, 5, 7, generator=gen) - 0.5) * 16.0).requires_grad_(True)
    w = torch.randn(2, N, 5, 7, generator=gen)
    cdf = O.bit_estimator(x, sd)
    np.testing.assert_allclose(cdf.detach().numpy(), g["cdf"], rtol=1e-6, atol=1e-7)
    f1 = O.bitparm(x, sd["bitEstimator.f1.h"], sd["bitEstimator.f1.b"], sd["bitEstimator.f1.a"])
    np.testing.assert_allclose(f1.detach().numpy(), g["f1"], rtol=1e-6, atol=1e-7)
    (cdf * w).sum().backward()
    np.testing.assert_allclose(x.grad.numpy(), g["dx"], rtol=1e-5, atol=1e-7)
    for k, p in sd.items():
        np.testing.assert_allclose(p.grad.numpy(), g["d/" + k[len("bitEstimator."):]], rtol=1e-4, atol=1e-6, err_msg=k)
