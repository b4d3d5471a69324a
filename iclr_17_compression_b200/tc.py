"""Host side of the tensor-core (tcgen05) inference path: packed weights, NHWC bf16 (hi, lo)
activations and fused layer launches through libb2c's C ABI (include/b2c.h, b2c_tc_*).
Inference only (no autograd); out_channel_N must be 128."""
import ctypes
import os

import torch

from . import _lib
from ._lib import TC_CONV1, TC_CONV5S2, TC_DECONV3, TC_DECONV3X, TC_DECONV5S2, TcLayer  # noqa: F401
from .functional import _timed

_vp = ctypes.c_void_p


def _p(t):
    return _vp(t.data_ptr()) if t is not None else _vp(0)


def _stream():
    return _vp(torch.cuda.current_stream().cuda_stream)


def deconv3_kind(channels):
    """Layer kind of the last synthesis layer (synthesis_17.py:23): the x-folded persistent kernel (DECONV3X: N = 112,
    3 MMAs per vertical tap and 16-channel slice) - B2C_DECONV3X=0 selects the 3x3-shift form (DECONV3) for A/B
    measurements.  The two kinds have different packed-weight layouts."""
    if channels in (128, 192, 256) and os.environ.get("B2C_DECONV3X", "1") != "0":
        return TC_DECONV3X
    return TC_DECONV3


def pack_weights(kind, w, need_lo):
    """state_dict weight (fp32 OIHW / IOHW) -> (hi, lo) bf16 tap-major panels."""
    lib = _lib.load()
    C = max(w.shape[0], w.shape[1])               # conv1 [C,3,9,9] / deconv3 [C,3,9,9] / 5x5 [C,C,5,5]
    n = lib.b2c_tc_packed_weight_elems(kind, C)
    if n == 0:
        raise _lib.B2CError(f"tensor-core kernels support 128, 192 or 256 channels, got {C}")
    w = w.detach().contiguous().float()
    hi = torch.empty((n,), device=w.device, dtype=torch.bfloat16)
    lo = torch.empty((n,), device=w.device, dtype=torch.bfloat16) if need_lo else None
    with torch.cuda.device(w.device):
        _lib.check(lib.b2c_tc_pack_weights(kind, C, _p(w), _p(hi), _p(lo), _stream()), "tc_pack_weights")
    return hi, lo


def pack_many(jobs):
    """Several packing jobs in ONE launch (b2c_tc_pack_many).  jobs: list of ("w", kind, weight) |
    ("gdn", beta, gamma) | ("gdn_t", beta, gamma); 128 channels, plain bf16 (no lo planes).  Returns the packs in the
    formats of pack_weights / pack_gdn / pack_gdn_t: (hi, None), (gamma, None, beta_eff)."""
    lib = _lib.load()
    outs, arr, keep = [], (_lib.TcPackJob * len(jobs))(), []
    dev = jobs[0][2].device
    for i, job in enumerate(jobs):
        J = arr[i]
        J.channels = 128
        if job[0] == "w":
            _, kind, w = job
            n = lib.b2c_tc_packed_weight_elems(kind, 128)
            w = w.detach().contiguous().float()
            hi = torch.empty((n,), device=dev, dtype=torch.bfloat16)
            J.kind, J.src, J.dst_hi = kind, w.data_ptr(), hi.data_ptr()
            keep.append(w)
            outs.append((hi, None))
        else:
            _, beta, gamma = job
            beta, gamma = beta.detach().contiguous().float(), gamma.detach().contiguous().float()
            g = torch.empty((128, 128), device=dev, dtype=torch.bfloat16)
            be = torch.empty((128,), device=dev, dtype=torch.float32)
            J.kind = _lib.TC_PACK_GDN if job[0] == "gdn" else _lib.TC_PACK_GDN_T
            J.src, J.src2, J.dst_hi, J.beta_eff = beta.data_ptr(), gamma.data_ptr(), g.data_ptr(), be.data_ptr()
            keep += [beta, gamma]
            outs.append((g, None, be))
    with torch.cuda.device(dev):
        _lib.check(lib.b2c_tc_pack_many(ctypes.cast(arr, ctypes.c_void_p), len(jobs), _stream()), "tc_pack_many")
    return outs


def pack_gdn(beta, gamma, need_lo):
    lib = _lib.load()
    beta, gamma = beta.detach().contiguous().float(), gamma.detach().contiguous().float()
    C = beta.numel()
    if C not in (128, 192, 256) or gamma.numel() != C * C:
        raise _lib.B2CError("tensor-core GDN needs 128, 192 or 256 channels")
    ghi = torch.empty((C, C), device=beta.device, dtype=torch.bfloat16)
    glo = torch.empty((C, C), device=beta.device, dtype=torch.bfloat16) if need_lo else None
    be = torch.empty((C,), device=beta.device, dtype=torch.float32)
    with torch.cuda.device(beta.device):
        _lib.check(lib.b2c_tc_pack_gdn(C, _p(beta), _p(gamma), _p(ghi), _p(glo), _p(be), _stream()), "tc_pack_gdn")
    return ghi, glo, be


def nchw_to_nhwc(x, need_lo):
    """fp32 NCHW -> NHWC bf16 hi[,lo]"""
    B, C, H, W = x.shape
    x = x.contiguous()
    hi = torch.empty((B, H, W, C), device=x.device, dtype=torch.bfloat16)
    lo = torch.empty_like(hi) if need_lo else None
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().b2c_tc_nchw_to_nhwc(_p(x), _p(hi), _p(lo), B, C, H * W, _stream()), "tc_nchw_to_nhwc")
    return hi, lo


def nhwc_to_nchw(hi, lo=None):
    B, H, W, C = hi.shape
    out = torch.empty((B, C, H, W), device=hi.device, dtype=torch.float32)
    with torch.cuda.device(hi.device):
        _lib.check(_lib.load().b2c_tc_nhwc_to_nchw(_p(hi), _p(lo), _p(out), B, C, H * W, _stream()), "tc_nhwc_to_nchw")
    return out


def _flops_per_out_pixel(kind, C):
    return {TC_CONV1: 2 * 243 * C, TC_CONV5S2: 2 * 25 * C * C, TC_DECONV5S2: 2 * 25 * C * C / 4,
            TC_DECONV3: 2 * C * 243 / 16, TC_DECONV3X: 2 * C * 243 / 16}[kind]


def pack_gdn_t(beta, gamma):
    """gamma_eff^T (bf16) and beta_eff, for the GDN backward epilogue"""
    lib = _lib.load()
    beta, gamma = beta.detach().contiguous().float(), gamma.detach().contiguous().float()
    C = beta.numel()
    gt = torch.empty((C, C), device=beta.device, dtype=torch.bfloat16)
    be = torch.empty((C,), device=beta.device, dtype=torch.float32)
    with torch.cuda.device(beta.device):
        _lib.check(lib.b2c_tc_pack_gdn_t(C, _p(beta), _p(gamma), _p(gt), _p(be), _stream()), "tc_pack_gdn_t")
    return gt, None, be


def reduce_partial_sse(partial):
    """[B, tiles, 2] per-tile squared-error sums of a DECONV3 launch -> [B, 2] (fixed order)"""
    B, tiles, _ = partial.shape
    sse = torch.empty((B, 2), device=partial.device, dtype=torch.float32)
    with torch.cuda.device(partial.device):
        _lib.check(_lib.load().b2c_partial_reduce(_p(partial), _p(sse), B, tiles, 2, _stream()), "partial_reduce")
    return sse


def layer_run(kind, x_hi, x_lo, w, bias, gdn=None, gdn_mode=0, split3=False, out_f32_nchw=False, ref_image=None,
              second_plane=False, saved=None, ksplit=1, out=None, grad_scale=None, reduce_sse=True, dbias=None,
              latent=None):
    """One fused layer.  x_hi/x_lo: NHWC bf16 (CONV1: x_hi = fp32 NCHW image).  w = (w_hi, w_lo);
    gdn = (gamma_hi, gamma_lo, beta_eff) with gdn_mode 1 (GDN) / 2 (IGDN).
    Returns (out_hi, out_lo) NHWC bf16, or the fp32 NCHW tensor when out_f32_nchw.
    ``latent`` = (bits table, half range, list of the 11 Bitparm vectors, RangeFlag or None): the last analysis layer
    at inference with the quantizer and the entropy model in its epilogue - returns (z fp32 NCHW, z bf16 NHWC,
    image_bits [B])."""
    if kind == TC_CONV1:
        B, _, Hin, Win = x_hi.shape
        C = w[0].numel() // (4 * 64)
        Ho, Wo, Cout = Hin // 4, Win // 4, C
    else:
        B, Hin, Win, C = x_hi.shape
        if kind == TC_CONV5S2:
            Ho, Wo, Cout = (Hin - 1) // 2 + 1, (Win - 1) // 2 + 1, C
        elif kind == TC_DECONV5S2:
            Ho, Wo, Cout = Hin * 2, Win * 2, C
        else:
            Ho, Wo, Cout = Hin * 4, Win * 4, 3
    dev = x_hi.device
    if out_f32_nchw:
        if out is not None:                           # caller-provided (zero-filled when ksplit > 1)
            out_hi = out.view(B, Cout, Ho, Wo)
        elif ksplit > 1:
            out_hi = torch.zeros((B, Cout, Ho, Wo), device=dev, dtype=torch.float32)
        else:
            out_hi = torch.empty((B, Cout, Ho, Wo), device=dev, dtype=torch.float32)
        out_lo = None
    else:
        out_hi = torch.empty((B, Ho, Wo, Cout), device=dev, dtype=torch.bfloat16)
        out_lo = torch.empty_like(out_hi) if (split3 or second_plane) else None
    L = TcLayer()
    L.kind, L.split3, L.gdn, L.out_f32_nchw = kind, int(split3), int(gdn_mode), int(out_f32_nchw)
    L.B, L.Hin, L.Win = B, Hin, Win
    L.channels = C
    L.ksplit = ksplit
    L.in_hi, L.in_lo = x_hi.data_ptr(), (x_lo.data_ptr() if x_lo is not None else 0)
    L.w_hi, L.w_lo = w[0].data_ptr(), (w[1].data_ptr() if w[1] is not None else 0)
    L.bias = bias.data_ptr() if bias is not None else 0
    if gdn_mode:
        L.gamma_hi, L.gamma_lo, L.beta_eff = gdn[0].data_ptr(), (gdn[1].data_ptr() if gdn[1] is not None else 0), \
            gdn[2].data_ptr()
    L.out_hi, L.out_lo = out_hi.data_ptr(), (out_lo.data_ptr() if out_lo is not None else 0)
    if saved is not None:
        L.saved_y, L.saved_n = saved[0].data_ptr(), saved[1].data_ptr()
    lat = None
    if latent is not None:                         # CONV5S2 -> z = rint(y), its bf16 NHWC copy and the bits per image
        lib = _lib.load()
        lut, half_range, params, flag = latent
        slots = lib.b2c_tc_tiles_per_image(kind, Hin, Win) * 8
        zh = torch.empty((B, Ho, Wo, Cout), device=dev, dtype=torch.bfloat16)
        bits_partial = torch.empty((B, slots), device=dev, dtype=torch.float32)
        image_bits = torch.empty((B,), device=dev, dtype=torch.float32)
        pp, keep = _lib.ptr_array(params)
        L.lut, L.lut_half_range, L.bit_params = lut.data_ptr(), int(half_range), ctypes.cast(pp, ctypes.c_void_p).value
        L.z_nhwc_out, L.bits_partial = zh.data_ptr(), bits_partial.data_ptr()
        L.range_flag = flag.ptr().value if flag is not None else 0
        lat = (zh, bits_partial, image_bits, slots, keep)
    sse = None
    if ref_image is not None:                      # DECONV3: fused clamp + squared errors
        lib = _lib.load()
        tiles = lib.b2c_tc_tiles_per_image(kind, Hin, Win)
        partial = torch.empty((B, tiles, 2), device=dev, dtype=torch.float32)
        sse = torch.empty((B, 2), device=dev, dtype=torch.float32)
        L.ref_image, L.sse_partial = ref_image.data_ptr(), partial.data_ptr()
        if grad_scale is not None:                 # training: distortion backward fused into the same epilogue
            grad_out = torch.empty_like(out_hi)
            L.grad_out, L.grad_scale = grad_out.data_ptr(), float(grad_scale)
            if dbias is not None:                  # DECONV3X: the bias gradient (sum of grad_out) from the same epilogue
                L.dbias_out = dbias.data_ptr()
    flops = int(_flops_per_out_pixel(kind, C) * B * Ho * Wo) + (2 * C * C * B * Ho * Wo if gdn_mode else 0)
    name = f"tc_layer[kind{kind},{B}x{Hin}x{Win},split{int(split3)},gdn{gdn_mode}]"
    with torch.cuda.device(dev), _timed(name, flops, 0):
        _lib.check(_lib.load().b2c_tc_layer_run(ctypes.byref(L), _stream()), "tc_layer_run")
        if sse is not None and reduce_sse:
            _lib.check(_lib.load().b2c_partial_reduce(_p(partial), _p(sse), B, tiles, 2, _stream()), "partial_reduce")
        if lat is not None and reduce_sse:
            _lib.check(_lib.load().b2c_partial_reduce(_p(lat[1]), _p(lat[2]), B, lat[3], 1, _stream()), "partial_reduce")
    if lat is not None:         # (reduce_sse=False: the caller reduces the bit sums, e.g. functional.infer_stats)
        return (out_hi, lat[0], lat[2]) if reduce_sse else (out_hi, lat[0], lat[2], lat[1])
    if sse is not None:
        if grad_scale is not None:
            return out_hi, (sse if reduce_sse else partial), grad_out
        return out_hi, (sse if reduce_sse else partial)
    return out_hi if out_f32_nchw else (out_hi, out_lo)


def wgrad(small, big, k, stride, pad, tmp=None, out=None, deterministic=False):
    """dW[m][n][ky][kx] = sum small[.,ys,xs,m] * big[., stride*ys-pad+ky, stride*xs-pad+kx, n]; NHWC bf16 in, fp32
    [128,128,k,k] out.

    Default: the CTAs accumulate with red.global.add into ``tmp`` (zero-filled fp32 [k,k,128,128], tap-major partial
    layout), and the result is transposed into ``out`` (or a new tensor).  ``deterministic=True``: per-CTA partial
    tiles + a fixed-order second kernel that ADDS into ``out`` (zero-filled) in the final layout - bit-reproducible,
    but measured slower at the training shapes (5x5: 54 vs 25 us, 1x1: 22 vs 10 us), so it is opt-in."""
    B, hs, ws, _ = small.shape
    _, hb, wb, _ = big.shape
    lib = _lib.load()
    flops = 2 * B * hs * ws * 128 * 128 * k * k
    name = f"tc_wgrad[{B}x{hs}x{ws},k{k}s{stride}]"
    if deterministic:
        dW = out.view(128, 128, k, k) if out is not None else torch.zeros((128, 128, k, k), device=small.device,
                                                                         dtype=torch.float32)
        nbytes = lib.b2c_tc_wgrad_workspace_bytes(B, hs, ws, k)
        work = torch.empty((nbytes,), device=small.device, dtype=torch.uint8)
        with torch.cuda.device(small.device), _timed(name, flops, 0):
            _lib.check(lib.b2c_tc_wgrad(_p(small), _p(big), _p(dW), B, hs, ws, hb, wb, k, stride, pad, _p(work), nbytes,
                                        _stream()), "tc_wgrad")
        return dW
    dW = tmp if tmp is not None else torch.zeros((k, k, 128, 128), device=small.device, dtype=torch.float32)
    with torch.cuda.device(small.device), _timed(name, flops, 0):
        _lib.check(lib.b2c_tc_wgrad(_p(small), _p(big), _p(dW), B, hs, ws, hb, wb, k, stride, pad, _vp(0), 0, _stream()),
                   "tc_wgrad")
    if k == 1:
        return dW.view(128, 128, 1, 1)
    res = out if out is not None else torch.empty((128, 128, k, k), device=small.device, dtype=torch.float32)
    with torch.cuda.device(small.device):          # tap-major accumulator -> the parameter's layout
        _lib.check(lib.b2c_tc_wgrad_untap(_p(dW), _p(res), k * k, 0, _stream()), "tc_wgrad_untap")
    return res


def wgrad9(small, image, out=None):
    """9x9/4 weight gradient: small NHWC bf16 [B,h,w,128], image fp32 NCHW [B,3,4h,4w] -> fp32 [128,3,9,9]
    (added to ``out``, which must be zero-filled).  Per-CTA partial tiles + fixed-order reduction (deterministic)."""
    B, h, w, _ = small.shape
    lib = _lib.load()
    dW = out if out is not None else torch.zeros((128, 3, 9, 9), device=small.device, dtype=torch.float32)
    nbytes = lib.b2c_tc_wgrad9_workspace_bytes(B, h, w)
    ws = torch.empty((nbytes,), device=small.device, dtype=torch.uint8)
    with torch.cuda.device(small.device), _timed(f"tc_wgrad9[{B}x{h}x{w}]", 2 * B * h * w * 128 * 243, 0):
        _lib.check(lib.b2c_tc_wgrad9(_p(small), _p(image), _p(dW), B, h, w, _p(ws), nbytes, _stream()), "tc_wgrad9")
    return dW
