"""Tensor-core (tcgen05) layers vs the exact fp32 operators of the same library and torch, op by op."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _maxabs(a, b):
    return (a.double().cpu() - b.double().cpu()).abs().max().item()


@pytest.fixture(scope="module")
def tc():
    import iclr_17_compression_b200.tc as t
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return t


def _gdn_params(g):
    from oracle import balle17_oracle as O
    beta = torch.sqrt(0.5 + torch.rand(128, generator=g))
    gamma = torch.sqrt(0.1 * torch.eye(128) + 0.02 * torch.rand(128, 128, generator=g) + O.PEDESTAL)
    beta[:3] = O.BETA_BOUND * 0.5
    gamma.view(-1)[5:40:7] = O.GAMMA_BOUND * 0.25
    return beta, gamma


def test_layout_roundtrip(tc):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 128, 9, 13, generator=g).cuda()
    hi, lo = tc.nchw_to_nhwc(x, True)
    assert _maxabs(tc.nhwc_to_nchw(hi, lo), x) < 2e-5
    assert torch.equal(hi, x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16))


@pytest.mark.parametrize("shape", [(1, 16, 32), (2, 32, 48), (1, 24, 40), (1, 128, 192)])
def test_conv5s2_split3_matches_fp32(tc, shape):
    B, H, W = shape
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, 128, H, W, generator=g).cuda()
    w = (torch.randn(128, 128, 5, 5, generator=g) * 0.03).cuda()
    ref = F.conv2d(x.double(), w.double(), None, stride=2, padding=2)
    xh, xl = tc.nchw_to_nhwc(x, True)
    wp = tc.pack_weights(tc.TC_CONV5S2, w, True)
    y = tc.layer_run(tc.TC_CONV5S2, xh, xl, wp, None, split3=True, out_f32_nchw=True)
    assert y.shape == ref.shape
    assert _rel(y, ref) < 3e-5, _rel(y, ref)
    assert _maxabs(y, ref) < 1e-4 * ref.abs().max().item()


@pytest.mark.parametrize("inverse", [False])
def test_conv5s2_fused_gdn(tc, inverse):
    from oracle import balle17_oracle as O
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 128, 32, 48, generator=g).cuda()
    w = (torch.randn(128, 128, 5, 5, generator=g) * 0.03).cuda()
    b = torch.randn(128, generator=g).cuda()
    beta, gamma = _gdn_params(g)
    ref = O.gdn(F.conv2d(x.double().cpu(), w.double().cpu(), b.double().cpu(), stride=2, padding=2),
                beta.double(), gamma.double(), inverse)
    xh, xl = tc.nchw_to_nhwc(x, True)
    wp = tc.pack_weights(tc.TC_CONV5S2, w, True)
    gp = tc.pack_gdn(beta.cuda(), gamma.cuda(), True)
    yh, yl = tc.layer_run(tc.TC_CONV5S2, xh, xl, wp, b, gdn=gp, gdn_mode=1, split3=True)
    y = tc.nhwc_to_nchw(yh, yl)
    assert _rel(y, ref) < 5e-5, _rel(y, ref)


@pytest.mark.parametrize("shape", [(1, 64, 96), (2, 128, 64), (1, 48, 80)])
def test_conv1_split3_fused_gdn(tc, shape):
    from oracle import balle17_oracle as O
    B, H, W = shape
    g = torch.Generator().manual_seed(3)
    x = torch.rand(B, 3, H, W, generator=g).cuda()
    w = (torch.randn(128, 3, 9, 9, generator=g) * 0.2).cuda()
    b = torch.randn(128, generator=g).cuda()
    beta, gamma = _gdn_params(g)
    conv = F.conv2d(x.double().cpu(), w.double().cpu(), b.double().cpu(), stride=4, padding=4)
    wp = tc.pack_weights(tc.TC_CONV1, w, True)
    y0 = tc.layer_run(tc.TC_CONV1, x, None, wp, b, split3=True, out_f32_nchw=True)
    assert _rel(y0, conv) < 3e-5, _rel(y0, conv)
    gp = tc.pack_gdn(beta.cuda(), gamma.cuda(), True)
    yh, yl = tc.layer_run(tc.TC_CONV1, x, None, wp, b, gdn=gp, gdn_mode=1, split3=True)
    ref = O.gdn(conv, beta.double(), gamma.double(), False)
    assert _rel(tc.nhwc_to_nchw(yh, yl), ref) < 5e-5


@pytest.mark.parametrize("shape", [(1, 8, 16), (2, 32, 48), (1, 12, 20)])
def test_deconv5s2_bf16_fused_igdn(tc, shape):
    from oracle import balle17_oracle as O
    B, H, W = shape
    g = torch.Generator().manual_seed(4)
    x = torch.randn(B, 128, H, W, generator=g).cuda()
    w = (torch.randn(128, 128, 5, 5, generator=g) * 0.03).cuda()
    b = torch.randn(128, generator=g).cuda()
    beta, gamma = _gdn_params(g)
    xh, _ = tc.nchw_to_nhwc(x, False)
    xq = xh.float().permute(0, 3, 1, 2).double().cpu()              # the bf16-rounded input the kernel sees
    wq = w.to(torch.bfloat16).double().cpu()
    dec = F.conv_transpose2d(xq, wq, b.double().cpu(), stride=2, padding=2, output_padding=1)
    wp = tc.pack_weights(tc.TC_DECONV5S2, w, False)
    y0 = tc.layer_run(tc.TC_DECONV5S2, xh, None, wp, b, split3=False, out_f32_nchw=True)
    assert y0.shape == dec.shape
    assert _rel(y0, dec) < 2e-5, _rel(y0, dec)                       # fp32 accumulation of exact bf16 products
    gp = tc.pack_gdn(beta.cuda(), gamma.cuda(), False)
    yh, _ = tc.layer_run(tc.TC_DECONV5S2, xh, None, wp, b, gdn=gp, gdn_mode=2, split3=False)
    ref = O.gdn(dec, beta.double(), gamma.double(), True)
    assert _rel(tc.nhwc_to_nchw(yh), ref) < 6e-3                     # plain-bf16 GDN operands + bf16 output


@pytest.mark.parametrize("shape", [(1, 16, 32), (2, 32, 48), (1, 20, 24)])
def test_deconv3_bf16(tc, shape):
    B, H, W = shape
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, 128, H, W, generator=g).cuda()
    w = (torch.randn(128, 3, 9, 9, generator=g) * 0.05).cuda()
    b = torch.randn(3, generator=g).cuda()
    xh, _ = tc.nchw_to_nhwc(x, False)
    xq = xh.float().permute(0, 3, 1, 2).double().cpu()
    wq = w.to(torch.bfloat16).double().cpu()
    ref = F.conv_transpose2d(xq, wq, b.double().cpu(), stride=4, padding=4, output_padding=3)
    wp = tc.pack_weights(tc.TC_DECONV3, w, False)
    y = tc.layer_run(tc.TC_DECONV3, xh, None, wp, b, split3=False, out_f32_nchw=True)
    assert y.shape == ref.shape
    assert _rel(y, ref) < 2e-5, _rel(y, ref)


@pytest.mark.parametrize("shape", [(1, 16, 32), (2, 32, 48), (1, 20, 24), (1, 8, 14), (3, 9, 29), (1, 40, 70)])
def test_deconv3x_bf16(tc, shape):
    """DECONV3X (horizontal shifts folded into N, summed by warp shuffles; tiles advance by 14 columns): same values
    as the transposed convolution of the bf16-rounded operands, plain and with the fused clamp + squared errors +
    distortion gradient; widths that are / are not multiples of 14 and 16, ragged heights."""
    B, H, W = shape
    g = torch.Generator().manual_seed(15)
    x = torch.randn(B, 128, H, W, generator=g).cuda()
    w = (torch.randn(128, 3, 9, 9, generator=g) * 0.05).cuda()
    b = torch.randn(3, generator=g).cuda()
    img = torch.rand(B, 3, 4 * H, 4 * W, generator=g).cuda()
    xh, _ = tc.nchw_to_nhwc(x, False)
    xq = xh.float().permute(0, 3, 1, 2).double().cpu()
    wq = w.to(torch.bfloat16).double().cpu()
    ref = F.conv_transpose2d(xq, wq, b.double().cpu(), stride=4, padding=4, output_padding=3)
    wp = tc.pack_weights(tc.TC_DECONV3X, w, False)
    y = tc.layer_run(tc.TC_DECONV3X, xh, None, wp, b, split3=False, out_f32_nchw=True)
    assert y.shape == ref.shape
    assert _rel(y, ref) < 2e-5, _rel(y, ref)
    clipped, sse, g_recon = tc.layer_run(tc.TC_DECONV3X, xh, None, wp, b, out_f32_nchw=True, ref_image=img,
                                         grad_scale=0.25)
    xi = img.double().cpu()
    assert torch.allclose(clipped.double().cpu(), ref.clamp(0, 1), atol=2e-5)
    sse_ref = torch.stack([((ref - xi) ** 2).flatten(1).sum(1), ((ref.clamp(0, 1) - xi) ** 2).flatten(1).sum(1)], 1)
    assert torch.allclose(sse.double().cpu(), sse_ref, rtol=1e-4), (sse, sse_ref)
    assert _rel(g_recon, 0.25 * (ref - xi)) < 2e-5
    # the 3x3-shift form computes the same layer
    y3 = tc.layer_run(tc.TC_DECONV3, xh, None, tc.pack_weights(tc.TC_DECONV3, w, False), b, out_f32_nchw=True)
    assert _rel(y, y3.double().cpu()) < 1e-5


@pytest.mark.parametrize("C", [192, 256])
@pytest.mark.parametrize("shape", [(1, 16, 32), (2, 24, 30), (1, 9, 29)])
def test_deconv3x_wide(tc, C, shape):
    """DECONV3X at the wider latent widths of BASELINE config 5 (model.py:39): 3 / 4 k-blocks, reference image through
    the TMA ring (192) or per-thread loads (256: the ring does not fit beside 168 KB of resident weights)."""
    B, H, W = shape
    g = torch.Generator().manual_seed(17)
    x = torch.randn(B, C, H, W, generator=g).cuda()
    w = (torch.randn(C, 3, 9, 9, generator=g) * 0.05).cuda()
    b = torch.randn(3, generator=g).cuda()
    img = torch.rand(B, 3, 4 * H, 4 * W, generator=g).cuda()
    xh, _ = tc.nchw_to_nhwc(x, False)
    xq = xh.float().permute(0, 3, 1, 2).double().cpu()
    wq = w.to(torch.bfloat16).double().cpu()
    ref = F.conv_transpose2d(xq, wq, b.double().cpu(), stride=4, padding=4, output_padding=3)
    wp = tc.pack_weights(tc.TC_DECONV3X, w, False)
    y = tc.layer_run(tc.TC_DECONV3X, xh, None, wp, b, out_f32_nchw=True)
    assert _rel(y, ref) < 2e-5, _rel(y, ref)
    clipped, sse = tc.layer_run(tc.TC_DECONV3X, xh, None, wp, b, out_f32_nchw=True, ref_image=img)
    xi = img.double().cpu()
    assert torch.allclose(clipped.double().cpu(), ref.clamp(0, 1), atol=2e-5)
    sse_ref = torch.stack([((ref - xi) ** 2).flatten(1).sum(1), ((ref.clamp(0, 1) - xi) ** 2).flatten(1).sum(1)], 1)
    assert torch.allclose(sse.double().cpu(), sse_ref, rtol=1e-4), (sse, sse_ref)


def test_deconv3x_full_size_properties(tc):
    """BASELINE configs[1] size (24 x 128 x 192 x 128 -> 24 x 3 x 512 x 768): the x-folded kernel against the independent
    3x3-shift kernel, linearity in the input, and the fused squared errors against a sum over the stored output."""
    g = torch.Generator().manual_seed(16)
    B, H, W = 24, 128, 192
    x = torch.randn(B, H, W, 128, generator=g).cuda().to(torch.bfloat16)
    w = (torch.randn(128, 3, 9, 9, generator=g) * 0.05).cuda()
    b = torch.randn(3, generator=g).cuda()
    wx, w3 = tc.pack_weights(tc.TC_DECONV3X, w, False), tc.pack_weights(tc.TC_DECONV3, w, False)
    y = tc.layer_run(tc.TC_DECONV3X, x, None, wx, b, out_f32_nchw=True)
    y3 = tc.layer_run(tc.TC_DECONV3, x, None, w3, b, out_f32_nchw=True)
    assert _rel(y, y3) < 1e-5, _rel(y, y3)
    # linearity: f(2x) - bias = 2 (f(x) - bias)   (2x is exact in bf16)
    y2 = tc.layer_run(tc.TC_DECONV3X, x * 2, None, wx, b, out_f32_nchw=True)
    bb = b.view(1, 3, 1, 1)
    assert _rel(y2 - bb, 2 * (y - bb)) < 1e-6
    img = torch.rand(B, 3, 4 * H, 4 * W, generator=g).cuda()
    clipped, sse = tc.layer_run(tc.TC_DECONV3X, x, None, wx, b, out_f32_nchw=True, ref_image=img)
    assert torch.equal(clipped, y.clamp(0, 1))
    ref = torch.stack([((y.double() - img.double()) ** 2).flatten(1).sum(1),
                       ((clipped.double() - img.double()) ** 2).flatten(1).sum(1)], 1)
    assert torch.allclose(sse.double(), ref, rtol=2e-5), (sse[:2], ref[:2])


@pytest.mark.parametrize("shape", [(1, 16, 32), (2, 32, 32), (3, 24, 40)])
def test_wgrad_5x5s2(tc, shape):
    """tensor-core weight gradient (MN-major MMAs) vs autograd, conv and transposed-conv forms."""
    B, H, W = shape
    g = torch.Generator().manual_seed(6)
    x = torch.randn(B, 128, H, W, generator=g).cuda()
    go = torch.randn(B, 128, H // 2, W // 2, generator=g).cuda()
    xh, _ = tc.nchw_to_nhwc(x, False)
    gh, _ = tc.nchw_to_nhwc(go, False)
    xq = xh.float().permute(0, 3, 1, 2).double().cpu()
    gq = gh.float().permute(0, 3, 1, 2).double().cpu()
    w = torch.zeros(128, 128, 5, 5, dtype=torch.float64, requires_grad=True)
    (F.conv2d(xq, w, None, stride=2, padding=2) * gq).sum().backward()
    dW = tc.wgrad(gh, xh, 5, 2, 2)                      # conv: small = grad_out, big = input -> OIHW
    assert _rel(dW, w.grad) < 1e-5, _rel(dW, w.grad)
    wt = torch.zeros(128, 128, 5, 5, dtype=torch.float64, requires_grad=True)
    (F.conv_transpose2d(gq, wt, None, stride=2, padding=2, output_padding=1) * xq).sum().backward()
    dWt = tc.wgrad(gh, xh, 5, 2, 2)                     # convT with input gq, grad_out xq: small = input, big = grad_out
    assert _rel(dWt, wt.grad) < 1e-5
    # deterministic two-stage form: same values, bit-identical from run to run, accumulates into `out`
    d1 = tc.wgrad(gh, xh, 5, 2, 2, deterministic=True)
    d2 = tc.wgrad(gh, xh, 5, 2, 2, deterministic=True)
    assert torch.equal(d1, d2) and _rel(d1, w.grad) < 1e-5
    acc = torch.full((128, 128, 5, 5), 0.5, device="cuda")
    tc.wgrad(gh, xh, 5, 2, 2, out=acc, deterministic=True)
    assert _rel(acc - 0.5, w.grad) < 1e-4


def test_wgrad_1x1_gamma_form(tc):
    g = torch.Generator().manual_seed(7)
    a = torch.randn(2, 128, 16, 24, generator=g).cuda()
    b = torch.randn(2, 128, 16, 24, generator=g).cuda()
    ah, _ = tc.nchw_to_nhwc(a, False)
    bh, _ = tc.nchw_to_nhwc(b, False)
    ref = torch.einsum("nihw,njhw->ij", ah.float().permute(0, 3, 1, 2).double(), bh.float().permute(0, 3, 1, 2).double())
    d = tc.wgrad(ah, bh, 1, 1, 0).view(128, 128)
    assert _rel(d, ref) < 1e-5
    assert _rel(tc.wgrad(ah, bh, 1, 1, 0, deterministic=True).view(128, 128), ref) < 1e-5


def _bf(t):
    return t.to(torch.bfloat16).double()


def test_train_forward_saves_norm_plane(tc):
    """bf16 forward with fused GDN: out_lo receives n = beta + gamma x^2 (needed by the backward)."""
    from oracle import balle17_oracle as O
    g = torch.Generator().manual_seed(8)
    x = torch.randn(2, 128, 32, 48, generator=g).cuda()
    w = (torch.randn(128, 128, 5, 5, generator=g) * 0.03).cuda()
    b = torch.randn(128, generator=g).cuda()
    beta, gamma = _gdn_params(g)
    xh, _ = tc.nchw_to_nhwc(x, False)
    conv = F.conv2d(_bf(x).cpu(), _bf(w).cpu(), b.double().cpu(), stride=2, padding=2)
    be, ge = O.gdn_effective_params(beta.double(), gamma.double())
    n_ref = F.conv2d(conv ** 2, ge.view(128, 128, 1, 1), be)
    y_ref = conv / n_ref.sqrt()
    wp = tc.pack_weights(tc.TC_CONV5S2, w, False)
    gp = tc.pack_gdn(beta.cuda(), gamma.cuda(), False)
    yh, nh = tc.layer_run(tc.TC_CONV5S2, xh, None, wp, b, gdn=gp, gdn_mode=1, split3=False, second_plane=True)
    assert _rel(tc.nhwc_to_nchw(yh), y_ref) < 6e-3
    assert _rel(tc.nhwc_to_nchw(nh), n_ref) < 6e-3


@pytest.mark.parametrize("kind_name,inverse", [("deconv", False), ("conv", True), ("conv1", True)])
def test_dgrad_with_gdn_backward_epilogue(tc, kind_name, inverse):
    """dgrad conv + fused (I)GDN backward: out_hi = dL/dx (pre-GDN), out_lo = dL/dnorm."""
    from oracle import balle17_oracle as O
    g = torch.Generator().manual_seed(9)
    beta, gamma = _gdn_params(g)
    be, ge = O.gdn_effective_params(beta.double(), gamma.double())
    if kind_name == "deconv":      # dgrad of a 5x5/2 conv: conv_transpose of g with the conv's OIHW weight
        gin = torch.randn(2, 128, 16, 24, generator=g).cuda()
        w = (torch.randn(128, 128, 5, 5, generator=g) * 0.03).cuda()
        gy = F.conv_transpose2d(_bf(gin).cpu(), _bf(w).cpu(), None, stride=2, padding=2, output_padding=1)
        kind, xin = tc.TC_DECONV5S2, tc.nchw_to_nhwc(gin, False)[0]
    elif kind_name == "conv":      # dgrad of a 5x5/2 transposed conv: conv of g with the IOHW weight
        gin = torch.randn(2, 128, 32, 48, generator=g).cuda()
        w = (torch.randn(128, 128, 5, 5, generator=g) * 0.03).cuda()
        gy = F.conv2d(_bf(gin).cpu(), _bf(w).cpu(), None, stride=2, padding=2)
        kind, xin = tc.TC_CONV5S2, tc.nchw_to_nhwc(gin, False)[0]
    else:                          # dgrad of deconv3 (9x9/4 transposed, 128->3): conv 9x9/4 of the fp32 image grad
        gin = torch.randn(2, 3, 64, 96, generator=g).cuda()
        w = (torch.randn(128, 3, 9, 9, generator=g) * 0.1).cuda()
        gy = F.conv2d(_bf(gin).cpu(), _bf(w).cpu(), None, stride=4, padding=4)
        kind, xin = tc.TC_CONV1, gin
    # a consistent forward state (x -> n, y), stored as bf16 like the training forward does
    x = torch.randn(gy.shape, generator=g, dtype=torch.float64)
    n = F.conv2d(x ** 2, ge.view(128, 128, 1, 1), be)
    y = x * n.sqrt() if inverse else x / n.sqrt()
    yq, nq = _bf(y), _bf(n)
    yh, nh = tc.nchw_to_nhwc(yq.float().cuda(), False)[0], tc.nchw_to_nhwc(nq.float().cuda(), False)[0]
    # reference backward from the stored (y, n)
    s = nq.sqrt()
    xq = yq / s if inverse else yq * s
    dn = (0.5 if inverse else -0.5) * gy * yq / nq
    t = F.conv2d(dn, ge.t().contiguous().view(128, 128, 1, 1))
    dx = (gy * s if inverse else gy / s) + 2 * xq * t
    wp = tc.pack_weights(kind, w, False)
    gp = tc.pack_gdn_t(beta.cuda(), gamma.cuda())
    dxh, dnh = tc.layer_run(kind, xin, None, wp, None, gdn=gp, gdn_mode=4 if inverse else 3, split3=False,
                            second_plane=True, saved=(yh, nh))
    assert _rel(tc.nhwc_to_nchw(dnh), dn) < 8e-3, _rel(tc.nhwc_to_nchw(dnh), dn)
    assert _rel(tc.nhwc_to_nchw(dxh), dx) < 8e-3, _rel(tc.nhwc_to_nchw(dxh), dx)


@pytest.mark.parametrize("shape", [(1, 16, 32), (2, 64, 64), (1, 24, 40)])
def test_wgrad9_im2col(tc, shape):
    B, h, w = shape
    g = torch.Generator().manual_seed(10)
    s = torch.randn(B, 128, h, w, generator=g).cuda()
    img = torch.randn(B, 3, 4 * h, 4 * w, generator=g).cuda()
    sh, _ = tc.nchw_to_nhwc(s, False)
    wt = torch.zeros(128, 3, 9, 9, dtype=torch.float64, requires_grad=True)
    (F.conv2d(_bf(img).cpu(), wt, None, stride=4, padding=4) * _bf(s).cpu()).sum().backward()
    d = tc.wgrad9(sh, img)
    assert _rel(d, wt.grad) < 1e-5, _rel(d, wt.grad)


@pytest.mark.parametrize("ksplit", [2, 5, 25])
@pytest.mark.parametrize("shape", [(8, 32, 32), (1, 24, 40)])
def test_conv5s2_bf16_split_k(tc, shape, ksplit):
    """split-K form of the plain-bf16 5x5/2 conv (training: conv3 forward, deconv1 dgrad): the taps are spread over
    ksplit CTAs per tile that add into the zero-filled fp32 output; must match the single-CTA result and add the bias once."""
    B, H, W = shape
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, 128, H, W, generator=g).cuda()
    w = (torch.randn(128, 128, 5, 5, generator=g) * 0.03).cuda()
    b = torch.randn(128, generator=g).cuda()
    xh, _ = tc.nchw_to_nhwc(x, False)
    wp = tc.pack_weights(tc.TC_CONV5S2, w, False)
    y1 = tc.layer_run(tc.TC_CONV5S2, xh, None, wp, b, out_f32_nchw=True)
    yk = tc.layer_run(tc.TC_CONV5S2, xh, None, wp, b, out_f32_nchw=True, ksplit=ksplit)
    assert _rel(yk, y1) < 1e-5, _rel(yk, y1)
    ref = F.conv2d(xh.permute(0, 3, 1, 2).double(), w.to(torch.bfloat16).double(), b.double(), stride=2, padding=2)
    assert _rel(yk, ref) < 1e-4, _rel(yk, ref)
    base = torch.full_like(y1, 0.25)                                  # accumulate on top of existing contents
    y2 = tc.layer_run(tc.TC_CONV5S2, xh, None, wp, b, out_f32_nchw=True, ksplit=ksplit, out=base)
    assert _rel(y2 - 0.25, y1) < 1e-5


_CONV1_FORM_SCRIPT = r"""
import sys, torch, torch.nn.functional as F
sys.path.insert(0, %r)
import iclr_17_compression_b200.tc as tc
from oracle import balle17_oracle as O
torch.backends.cudnn.allow_tf32 = False
g = torch.Generator().manual_seed(3)
worst = 0.0
for (B, H, W) in [(1, 64, 96), (3, 48, 80), (1, 16, 16), (2, 272, 400)]:      # incl. odd tile counts / ragged borders
    x = torch.rand(B, 3, H, W, generator=g).cuda()
    w = (torch.randn(128, 3, 9, 9, generator=g) * 0.2).cuda()
    b = torch.randn(128, generator=g).cuda()
    beta = torch.sqrt(0.5 + torch.rand(128, generator=g))
    gamma = torch.sqrt(0.1 * torch.eye(128) + 0.02 * torch.rand(128, 128, generator=g) + O.PEDESTAL)
    conv = F.conv2d(x.double().cpu(), w.double().cpu(), b.double().cpu(), stride=4, padding=4)
    ref = O.gdn(conv, beta.double(), gamma.double(), False)
    wp = tc.pack_weights(tc.TC_CONV1, w, True)
    gp = tc.pack_gdn(beta.cuda(), gamma.cuda(), True)
    for rep in range(2):                                                       # second call: steady-state barriers / flags
        yh, yl = tc.layer_run(tc.TC_CONV1, x, None, wp, b, gdn=gp, gdn_mode=1, split3=True)
    y = tc.nhwc_to_nchw(yh, yl).double().cpu()
    worst = max(worst, ((y - ref).norm() / ref.norm()).item())
print("WORST", worst)
"""


@pytest.mark.parametrize("form", ["3", "2", "1", "0"])
def test_conv1_every_kernel_form(form):
    """B2C_TC_PERSIST_CONV1 selects the CONV1 + GDN1 kernel (3 = role-specialised with the MMA A operands in TMEM, the
    default; 2 = single worker group, operands in TMEM; 1 = operands in shared memory; 0 = per-tile kernel).  The switch
    is read once per process, so every form runs in its own interpreter, against the fp64 oracle."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, B2C_TC_PERSIST_CONV1=form)
    out = subprocess.run([sys.executable, "-c", _CONV1_FORM_SCRIPT % root], capture_output=True, text=True, env=env,
                         timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    worst = float(out.stdout.strip().split("WORST")[-1])
    assert worst < 5e-5, worst


def test_pack_many_matches_single_packs(tc):
    """b2c_tc_pack_many (one launch, 5x5 panels through shared memory) == the per-operand packing calls, bit for bit"""
    g = torch.Generator().manual_seed(21)
    w5 = (torch.randn(128, 128, 5, 5, generator=g) * 0.03).cuda()
    w9 = (torch.randn(128, 3, 9, 9, generator=g) * 0.05).cuda()
    beta, gamma = _gdn_params(g)
    beta, gamma = beta.cuda(), gamma.cuda()
    many = tc.pack_many([("w", tc.TC_CONV5S2, w5), ("w", tc.TC_DECONV5S2, w5), ("w", tc.TC_CONV1, w9),
                         ("w", tc.TC_DECONV3, w9), ("w", tc.TC_DECONV3X, w9), ("gdn", beta, gamma),
                         ("gdn_t", beta, gamma)])
    for got, kind, w in zip(many[:5], (tc.TC_CONV5S2, tc.TC_DECONV5S2, tc.TC_CONV1, tc.TC_DECONV3, tc.TC_DECONV3X),
                            (w5, w5, w9, w9, w9)):
        assert torch.equal(got[0], tc.pack_weights(kind, w, False)[0]), kind
    g1, g2 = tc.pack_gdn(beta, gamma, False), tc.pack_gdn_t(beta, gamma)
    assert torch.equal(many[5][0], g1[0]) and torch.equal(many[5][2], g1[2])
    assert torch.equal(many[6][0], g2[0]) and torch.equal(many[6][2], g2[2])
