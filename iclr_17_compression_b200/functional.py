"""autograd Functions over the libb2c C ABI (include/b2c.h): the exact-fp32 operator set of the
Balle-2017 hot path.  Tensors are fp32 NCHW CUDA tensors, as in the reference.  PyTorch is used
only for memory, streams and autograd bookkeeping; all arithmetic happens in libb2c kernels."""
import contextlib
import ctypes

import torch

from . import _lib

_vp = ctypes.c_void_p

# ------------------------------------------------------------------------------ op timing (bench.py)
_timer = None


class OpTimer:
    """CUDA-event pairs around every libb2c call on the launching stream (torch's current stream is
    the stream handed to the library), with the call's algorithmic FLOPs / bytes."""

    def __init__(self):
        self.records = []          # (name, flops, bytes, start_event, end_event)

    def summary(self):
        """name -> dict(calls, ms_total, ms_avg, flops, bytes); call after torch.cuda.synchronize()."""
        out = {}
        for name, fl, by, s, e in self.records:
            d = out.setdefault(name, {"calls": 0, "ms_total": 0.0, "flops": fl, "bytes": by})
            d["calls"] += 1
            d["ms_total"] += s.elapsed_time(e)
        for d in out.values():
            d["ms_avg"] = d["ms_total"] / d["calls"]
        return out


def start_timing():
    global _timer
    _timer = OpTimer()
    return _timer


def stop_timing():
    global _timer
    t, _timer = _timer, None
    return t


@contextlib.contextmanager
def _timed(name, flops=0, nbytes=0):
    if _timer is None:
        yield
        return
    s = torch.cuda.Event(enable_timing=True)
    e = torch.cuda.Event(enable_timing=True)
    s.record()
    yield
    e.record()
    _timer.records.append((name, flops, nbytes, s, e))


def _p(t):
    return _vp(t.data_ptr()) if t is not None else _vp(0)


def _stream():
    return _vp(torch.cuda.current_stream().cuda_stream)


def _prep(t, name):
    if not t.is_cuda:
        raise _lib.B2CError(f"{name} must be a CUDA tensor (this package has no CPU path)")
    if t.dtype != torch.float32:
        raise _lib.B2CError(f"{name} must be float32, got {t.dtype}")
    _lib.require_device(t.device.index)
    return t.contiguous()


# ------------------------------------------------------------------------------ raw calls
def _conv2d_raw(x, w, b, stride, pad):
    B, Cin, H, W = x.shape
    Cout, Cin2, k, k2 = w.shape
    if Cin2 != Cin or k != k2:
        raise _lib.B2CError(f"conv2d: weight {tuple(w.shape)} does not match input {tuple(x.shape)}")
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    y = torch.empty((B, Cout, Ho, Wo), device=x.device, dtype=torch.float32)
    name = f"conv2d_fwd[{B}x{Cin}x{H}x{W}->{Cout},k{k}s{stride}]"
    with torch.cuda.device(x.device), _timed(name, 2 * B * Ho * Wo * Cin * k * k * Cout,
                                             4 * (x.numel() + w.numel() + y.numel())):
        _lib.check(_lib.load().b2c_conv2d_fwd(_p(x), _p(w), _p(b), _p(y), B, Cin, H, W, Cout, k, stride, pad,
                                              _stream()), "conv2d_fwd")
    return y


def _convT_raw(x, w, b, stride, pad, out_pad):
    B, Cin, H, W = x.shape
    Cin2, Cout, k, k2 = w.shape
    if Cin2 != Cin or k != k2:
        raise _lib.B2CError(f"conv_transpose2d: weight {tuple(w.shape)} does not match input {tuple(x.shape)}")
    Ho, Wo = (H - 1) * stride - 2 * pad + k + out_pad, (W - 1) * stride - 2 * pad + k + out_pad
    y = torch.empty((B, Cout, Ho, Wo), device=x.device, dtype=torch.float32)
    name = f"conv_transpose2d_fwd[{B}x{Cin}x{H}x{W}->{Cout},k{k}s{stride}]"
    with torch.cuda.device(x.device), _timed(name, 2 * B * H * W * Cin * k * k * Cout,
                                             4 * (x.numel() + w.numel() + y.numel())):
        _lib.check(_lib.load().b2c_conv_transpose2d_fwd(_p(x), _p(w), _p(b), _p(y), B, Cin, H, W, Cout, k, stride,
                                                        pad, out_pad, _stream()), "conv_transpose2d_fwd")
    return y


def _wgrad_raw(small, big, k, stride, pad):
    B, Cs, hs, ws = small.shape
    _, Cb, hb, wb = big.shape
    dW = torch.zeros((Cs, Cb, k, k), device=small.device, dtype=torch.float32)
    name = f"conv_wgrad[{B}x{Cs}x{hs}x{ws}*{Cb}x{hb}x{wb},k{k}s{stride}]"
    with torch.cuda.device(small.device), _timed(name, 2 * B * hs * ws * Cs * Cb * k * k,
                                                 4 * (small.numel() + big.numel() + dW.numel())):
        _lib.check(_lib.load().b2c_conv_wgrad(_p(small), _p(big), _p(dW), B, Cs, hs, ws, Cb, hb, wb, k, stride, pad,
                                              _stream()), "conv_wgrad")
    return dW


def _channel_sum_raw(t, out=None):
    B, C = t.shape[0], t.shape[1]
    if out is None:
        out = torch.empty((C,), device=t.device, dtype=torch.float32)
    with torch.cuda.device(t.device):
        _lib.check(_lib.load().b2c_channel_sum(_p(t), _p(out), B, C, t[0, 0].numel(), _stream()), "channel_sum")
    return out


# ------------------------------------------------------------------------------ autograd
class _Conv2d(torch.autograd.Function):
    """nn.Conv2d (models/analysis_17.py:14,18,22)."""

    @staticmethod
    def forward(ctx, x, w, b, stride, pad):
        x, w = _prep(x, "input"), _prep(w, "weight")
        b = _prep(b, "bias") if b is not None else None
        ctx.save_for_backward(x, w)
        ctx.cfg = (stride, pad, b is not None)
        return _conv2d_raw(x, w, b, stride, pad)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        stride, pad, has_b = ctx.cfg
        g = g.contiguous()
        k = w.shape[2]
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            # dgrad(conv) = conv_transpose with the same weight memory (OIHW == IOHW of the transpose)
            out_pad = x.shape[2] - ((g.shape[2] - 1) * stride - 2 * pad + k)
            gx = _convT_raw(g, w, None, stride, pad, out_pad)
        if ctx.needs_input_grad[1]:
            gw = _wgrad_raw(g, x, k, stride, pad)
        if has_b and ctx.needs_input_grad[2]:
            gb = _channel_sum_raw(g)
        return gx, gw, gb, None, None


class _ConvTranspose2d(torch.autograd.Function):
    """nn.ConvTranspose2d (models/synthesis_17.py:15,19,23)."""

    @staticmethod
    def forward(ctx, x, w, b, stride, pad, out_pad):
        x, w = _prep(x, "input"), _prep(w, "weight")
        b = _prep(b, "bias") if b is not None else None
        ctx.save_for_backward(x, w)
        ctx.cfg = (stride, pad, b is not None)
        return _convT_raw(x, w, b, stride, pad, out_pad)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        stride, pad, has_b = ctx.cfg
        g = g.contiguous()
        k = w.shape[2]
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = _conv2d_raw(g, w, None, stride, pad)     # dgrad(convT) = conv with the same weight memory
        if ctx.needs_input_grad[1]:
            gw = _wgrad_raw(x, g, k, stride, pad)
        if has_b and ctx.needs_input_grad[2]:
            gb = _channel_sum_raw(g)
        return gx, gw, gb, None, None, None


class _GDN(torch.autograd.Function):
    """models/GDN.py:64-94 incl. the LowerBound reparametrisation (GDN.py:10-24,73-79)."""

    @staticmethod
    def forward(ctx, x, beta, gamma, inverse):
        x, beta, gamma = _prep(x, "input"), _prep(beta, "beta"), _prep(gamma, "gamma")
        B, C = x.shape[0], x.shape[1]
        if beta.numel() != C or gamma.numel() != C * C:
            raise _lib.B2CError(f"GDN: beta/gamma do not match {C} channels")
        HW = x[0, 0].numel()
        y = torch.empty_like(x)
        with torch.cuda.device(x.device), _timed(f"gdn_fwd[{B}x{C}x{HW}]", B * HW * C * (2 * C + 4),
                                                 8 * x.numel()):
            _lib.check(_lib.load().b2c_gdn_fwd(_p(x), _p(beta), _p(gamma), _p(y), B, C, HW, int(inverse), _stream()),
                       "gdn_fwd")
        ctx.save_for_backward(x, beta, gamma)
        ctx.inverse = bool(inverse)
        return y

    @staticmethod
    def backward(ctx, g):
        x, beta, gamma = ctx.saved_tensors
        g = g.contiguous()
        B, C = x.shape[0], x.shape[1]
        HW = x[0, 0].numel()
        lib = _lib.load()
        gx = torch.empty_like(x)
        dbeta = torch.empty_like(beta)
        dgamma = torch.empty_like(gamma)
        nbytes = lib.b2c_gdn_bwd_workspace_bytes(B, C, HW)
        ws = torch.empty((nbytes,), device=x.device, dtype=torch.uint8)
        with torch.cuda.device(x.device), _timed(f"gdn_bwd[{B}x{C}x{HW}]", 3 * B * HW * C * (2 * C + 4),
                                                 16 * x.numel()):
            _lib.check(lib.b2c_gdn_bwd(_p(x), _p(beta), _p(gamma), _p(g), _p(gx), _p(dbeta), _p(dgamma), B, C, HW,
                                       int(ctx.inverse), _p(ws), nbytes, _stream()), "gdn_bwd")
        return gx, dbeta, dgamma, None


class _QuantEntropy(torch.autograd.Function):
    """model.py:53-56 (quantizer) + model.py:71-73 (bits) + models/bitEstimator.py:20-42."""

    @staticmethod
    def forward(ctx, y, noise, *params):
        y = _prep(y, "latent")
        params = [_prep(p, "bit-estimator parameter") for p in params]
        assert len(params) == 11
        B, C = y.shape[0], y.shape[1]
        HW = y[0, 0].numel()
        for p in params:
            if p.numel() != C:
                raise _lib.B2CError("bit-estimator parameter does not match the latent's channel count")
        if noise is not None:
            noise = _prep(noise, "noise")
            if noise.shape != y.shape:
                raise _lib.B2CError(f"quantization noise {tuple(noise.shape)} does not match latent {tuple(y.shape)}")
        z = torch.empty_like(y)
        plane_bits = torch.empty((B, C), device=y.device, dtype=torch.float32)
        image_bits = torch.empty((B,), device=y.device, dtype=torch.float32)
        pp, keep = _lib.ptr_array(params)
        nb = (12 if noise is not None else 8) * y.numel()
        with torch.cuda.device(y.device), _timed(f"quant_entropy_fwd[{B}x{C}x{HW}]", 60 * y.numel(), nb):
            _lib.check(_lib.load().b2c_quant_entropy_fwd(_p(y), _p(noise), pp, _p(z), _p(plane_bits), _p(image_bits),
                                                         B, C, HW, _stream()), "quant_entropy_fwd")
        total_bits = image_bits.sum()              # model.py:73 sums over the whole batch
        ctx.save_for_backward(z, *params)
        ctx.training = noise is not None
        ctx.mark_non_differentiable(image_bits, plane_bits)
        return z, total_bits, image_bits, plane_bits

    @staticmethod
    def backward(ctx, gz, g_total, _g_image, _g_plane):
        z, *params = ctx.saved_tensors
        B, C = z.shape[0], z.shape[1]
        HW = z[0, 0].numel()
        dparams = [torch.zeros_like(p) for p in params]
        dz = None
        if g_total is not None:
            g_bits = g_total.reshape(1).contiguous().float()
            dz = torch.empty_like(z)
            pp, k1 = _lib.ptr_array(params)
            dp, k2 = _lib.ptr_array(dparams)
            with torch.cuda.device(z.device), _timed(f"entropy_bwd[{B}x{C}x{HW}]", 150 * z.numel(), 8 * z.numel()):
                _lib.check(_lib.load().b2c_entropy_bwd(_p(z), pp, _p(g_bits), _p(dz), dp, B, C, HW, _stream()),
                           "entropy_bwd")
        gy = None
        if ctx.needs_input_grad[0]:
            if ctx.training:                       # z = y + u  ->  dz/dy = 1
                gy = dz if gz is None else (gz if dz is None else gz + dz)
            else:                                  # z = round(y) -> zero gradient (torch.round)
                gy = torch.zeros_like(z)
        return (gy, None, *dparams)


class _Bitparm(torch.autograd.Function):
    """models/bitEstimator.py:20-25 as a module call: non-final ``t + tanh(t) * tanh(a)``, final ``sigmoid(t)`` with
    ``t = x * softplus(h) + b``; h, b, a are the [1,C,1,1] parameters, x is [B,C,H,W]."""

    @staticmethod
    def forward(ctx, x, h, b, a):
        x = _prep(x, "input")
        if x.dim() != 4:
            raise _lib.B2CError(f"Bitparm expects a [B,C,H,W] tensor, got {tuple(x.shape)}")
        B, C = x.shape[0], x.shape[1]
        HW = x[0, 0].numel()
        h, b = _prep(h, "h").reshape(-1), _prep(b, "b").reshape(-1)
        a = _prep(a, "a").reshape(-1) if a is not None else None
        if h.numel() != C or b.numel() != C or (a is not None and a.numel() != C):
            raise _lib.B2CError(f"Bitparm parameters do not match {C} channels")
        out = torch.empty_like(x)
        if x.numel():
            with torch.cuda.device(x.device), _timed(f"bitparm_fwd[{B}x{C}x{HW}]", 10 * x.numel(), 8 * x.numel()):
                _lib.check(_lib.load().b2c_bitparm_fwd(_p(x), _p(h), _p(b), _p(a), _p(out), B, C, HW, _stream()),
                           "bitparm_fwd")
        ctx.save_for_backward(x, h, b, a)
        return out

    @staticmethod
    def backward(ctx, g):
        x, h, b, a = ctx.saved_tensors
        B, C = x.shape[0], x.shape[1]
        HW = x[0, 0].numel()
        g = _prep(g, "grad")
        lib = _lib.load()
        dx = torch.empty_like(x)
        dh, db = torch.zeros_like(h), torch.zeros_like(b)
        da = torch.zeros_like(a) if a is not None else None
        if x.numel():
            nbytes = lib.b2c_bitparm_bwd_workspace_bytes(B, C)
            ws = torch.empty((nbytes,), device=x.device, dtype=torch.uint8)
            with torch.cuda.device(x.device), _timed(f"bitparm_bwd[{B}x{C}x{HW}]", 20 * x.numel(), 12 * x.numel()):
                _lib.check(lib.b2c_bitparm_bwd(_p(x), _p(h), _p(b), _p(a), _p(g), _p(dx), _p(dh), _p(db), _p(da), B, C, HW,
                                               _p(ws), nbytes, _stream()), "bitparm_bwd")
        shape = (1, C, 1, 1)
        return dx, dh.view(shape), db.view(shape), (da.view(shape) if da is not None else None)


def bitparm(x, h, b, a=None):
    """one Bitparm layer (bitEstimator.py:20-25); ``a=None`` selects the final (sigmoid) form.  Differentiable."""
    return _Bitparm.apply(x, h, b, a)


class _Distortion(torch.autograd.Function):
    """model.py:59 (clamp) + model.py:61 (squared error on the unclamped reconstruction)."""

    @staticmethod
    def forward(ctx, recon, x):
        recon, x = _prep(recon, "reconstruction"), _prep(x, "input image")
        if recon.shape != x.shape:
            raise _lib.B2CError(f"reconstruction {tuple(recon.shape)} does not match input {tuple(x.shape)}")
        B = x.shape[0]
        CHW = x[0].numel()
        lib = _lib.load()
        clipped = torch.empty_like(recon)
        sse = torch.empty((B, 2), device=x.device, dtype=torch.float32)
        nbytes = lib.b2c_clamp_sse_workspace_bytes(B, CHW)
        ws = torch.empty((nbytes,), device=x.device, dtype=torch.uint8)
        with torch.cuda.device(x.device), _timed(f"clamp_sse_fwd[{B}x{CHW}]", 8 * x.numel(), 12 * x.numel()):
            _lib.check(lib.b2c_clamp_sse_fwd(_p(recon), _p(x), _p(clipped), _p(sse), B, CHW, _p(ws), nbytes,
                                             _stream()), "clamp_sse_fwd")
        sse_total = sse[:, 0].sum()                # model.py:61 averages over the whole batch
        ctx.save_for_backward(recon, x)
        ctx.mark_non_differentiable(sse)
        return clipped, sse_total, sse

    @staticmethod
    def backward(ctx, g_clipped, g_total, _g_sse):
        recon, x = ctx.saved_tensors
        if g_total is None:
            g_s = torch.zeros((1,), device=x.device, dtype=torch.float32)
        else:
            g_s = g_total.reshape(1).contiguous().float()
        gc = g_clipped.contiguous() if g_clipped is not None else None
        g_recon = torch.empty_like(recon)
        with torch.cuda.device(x.device), _timed("clamp_sse_bwd", 4 * x.numel(), 16 * x.numel()):
            _lib.check(_lib.load().b2c_clamp_sse_bwd(_p(recon), _p(x), _p(gc), _p(g_s), _p(g_recon), recon.numel(),
                                                     _stream()), "clamp_sse_bwd")
        return g_recon, None


LUT_HALF_RANGE = 32      # table covers z in [-32, 32]; larger |z| are evaluated directly inside the kernel


def entropy_lut(params):
    """bits[c][z + 32] for integer z, the inference entropy table (same formula as the direct kernel); build once per
    weight version"""
    params = [_prep(p, "bit-estimator parameter") for p in params]
    C = params[0].numel()
    lut = torch.empty((C, 2 * LUT_HALF_RANGE + 1), device=params[0].device, dtype=torch.float32)
    pp, keep = _lib.ptr_array(params)
    with torch.cuda.device(lut.device):
        _lib.check(_lib.load().b2c_entropy_lut_build(pp, _p(lut), C, LUT_HALF_RANGE, _stream()), "entropy_lut_build")
    return lut


class RangeFlag:
    """Sticky 'some |z| > 256' flag of the inference quantizer, living in mapped pinned host memory: the kernel
    writes it through the device's view of the page, the host reads it WITHOUT synchronising.  Beyond 256 the NHWC
    bf16 latent the tensor-core synthesis transform consumes is inexact; callers check the flag wherever they
    synchronise anyway (HostPipeline.compress, bitstream.compress) and ImageCompressor re-checks it at the start of
    every tensor-core forward."""

    def __init__(self):
        self.host = torch.zeros((1,), dtype=torch.int32).pin_memory()

    def ptr(self):
        return _vp(self.host.data_ptr())          # unified addressing: pinned host pointers are valid on the device

    def raised(self):
        return bool(int(self.host[0]))

    def clear(self):
        self.host.zero_()


def quant_entropy_infer(y, params, lut=None, range_flag=None):
    """eval quantizer + bits for the tensor-core pipeline (no autograd): (z fp32 NCHW, z NHWC bf16, image_bits[B]).
    With ``lut`` (entropy_lut(params)) the table-driven HBM-bound kernel runs.  ``range_flag``: a RangeFlag."""
    y = _prep(y, "latent")
    params = [_prep(p, "bit-estimator parameter") for p in params]
    B, C, H, W = y.shape
    HW = H * W
    lib = _lib.load()
    z = torch.empty_like(y)
    zh = torch.empty((B, H, W, C), device=y.device, dtype=torch.bfloat16)
    image_bits = torch.empty((B,), device=y.device, dtype=torch.float32)
    nbytes = lib.b2c_quant_entropy_infer_workspace_bytes(B, HW)
    ws = torch.empty((nbytes,), device=y.device, dtype=torch.uint8)
    pp, keep = _lib.ptr_array(params)
    rf = range_flag.ptr() if range_flag is not None else _vp(0)
    with torch.cuda.device(y.device), _timed(f"quant_entropy_infer[{B}x{C}x{HW}]", 60 * y.numel(), 10 * y.numel()):
        if lut is not None and C % 8 == 0 and C <= 256 and HW % 4 == 0:
            _lib.check(lib.b2c_quant_entropy_infer_lut(_p(y), pp, _p(lut), LUT_HALF_RANGE, _p(z), _p(zh), _p(image_bits),
                                                       rf, B, C, HW, _p(ws), nbytes, _stream()), "quant_entropy_infer_lut")
        else:
            _lib.check(lib.b2c_quant_entropy_infer(_p(y), pp, _p(z), _p(zh), _p(image_bits), rf, B, C, HW, _p(ws),
                                                   nbytes, _stream()), "quant_entropy_infer")
    return z, zh, image_bits


_stats_tickets = {}


def infer_stats(sse_partial, image_bits, numel, pixels, bits_partial=None):
    """End of an eval forward in one launch (b2c_infer_stats; model.py:61,73,78): per-tile squared-error partial sums
    [B, tiles, 2] + per-image bits [B] -> (sse [B, 2], totals [4] = total bits, total squared error, mse, bpp).
    ``bits_partial`` [B, slots]: per-(tile, warp) bit sums of the fused quantizer epilogue - ``image_bits`` is then
    filled by this launch as well.
    The kernel finds its last block through a self-resetting ticket word; calls take the 64 words of a per-device pool in
    turn, so forwards in flight on different streams (or captured in different graphs) do not share one."""
    B, tiles, _ = sse_partial.shape
    dev = sse_partial.device
    pool = _stats_tickets.get(dev.index)
    if pool is None:
        if torch.cuda.is_current_stream_capturing():
            raise _lib.B2CError("infer_stats: first call on this device inside a CUDA-graph capture; run one eager "
                                "forward first")
        pool = _stats_tickets[dev.index] = [torch.zeros((64,), device=dev, dtype=torch.int32), 0]
    ticket = pool[0][pool[1] % 64:pool[1] % 64 + 1]
    pool[1] += 1
    sse = torch.empty((B, 2), device=dev, dtype=torch.float32)
    totals = torch.empty((4,), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev), _timed("infer_stats", 0, sse_partial.numel() * 4):
        _lib.check(_lib.load().b2c_infer_stats(_p(sse_partial), B, tiles, _p(image_bits), _p(bits_partial),
                                               bits_partial.shape[1] if bits_partial is not None else 0, float(numel),
                                               float(pixels), _p(sse), _p(totals), _p(ticket), _stream()), "infer_stats")
    return sse, totals


class _SigmoidBinarize(torch.autograd.Function):
    """models/analysis_17_new.py:35-37 + models/binarizer.py:7-27: (code, sigmoid) with the straight-through gradient"""

    @staticmethod
    def forward(ctx, y):
        y = _prep(y, "conv3 output")
        B, C = y.shape[0], y.shape[1]
        HW = y[0, 0].numel()
        sig, code = torch.empty_like(y), torch.empty_like(y)
        with torch.cuda.device(y.device), _timed(f"sigmoid_binarize_fwd[{B}x{C}x{HW}]", 8 * y.numel(), 12 * y.numel()):
            _lib.check(_lib.load().b2c_sigmoid_binarize_fwd(_p(y), _p(sig), _p(code), _vp(0), B, C, HW, _stream()),
                       "sigmoid_binarize_fwd")
        ctx.save_for_backward(sig)
        return code, sig

    @staticmethod
    def backward(ctx, g_code, g_sig):
        (sig,) = ctx.saved_tensors
        g_code = _prep(g_code, "grad") if g_code is not None else None
        g_sig = _prep(g_sig, "grad") if g_sig is not None else None
        gy = torch.empty_like(sig)
        with torch.cuda.device(sig.device):
            _lib.check(_lib.load().b2c_sigmoid_binarize_bwd(_p(sig), _p(g_code), _p(g_sig), _p(gy), sig.numel(), _stream()),
                       "sigmoid_binarize_bwd")
        return gy


def sigmoid_binarize(y):
    """(code, sigmoid(y)); differentiable (straight-through binarizer)"""
    return _SigmoidBinarize.apply(y)


def sigmoid_binarize_infer(y):
    """eval, no autograd: (code fp32 NCHW, sigmoid fp32 NCHW, code NHWC bf16 for the tensor-core decoder)"""
    y = _prep(y, "conv3 output")
    B, C, H, W = y.shape
    sig, code = torch.empty_like(y), torch.empty_like(y)
    ch = torch.empty((B, H, W, C), device=y.device, dtype=torch.bfloat16)
    with torch.cuda.device(y.device), _timed(f"sigmoid_binarize_fwd[{B}x{C}x{H * W}]", 8 * y.numel(), 14 * y.numel()):
        _lib.check(_lib.load().b2c_sigmoid_binarize_fwd(_p(y), _p(sig), _p(code), _p(ch), B, C, H * W, _stream()),
                   "sigmoid_binarize_fwd")
    return code, sig, ch


# ------------------------------------------------------------------------------ public ops
def conv2d(x, weight, bias, stride, padding):
    return _Conv2d.apply(x, weight, bias, stride, padding)


def conv_transpose2d(x, weight, bias, stride, padding, output_padding):
    return _ConvTranspose2d.apply(x, weight, bias, stride, padding, output_padding)


def gdn(x, beta, gamma, inverse=False):
    return _GDN.apply(x, beta, gamma, inverse)


def quant_entropy(y, noise, params):
    """returns (z, total_bits, image_bits[B], plane_bits[B,C]); noise=None selects the eval (round)
    quantizer.  Only z and total_bits are differentiable."""
    return _QuantEntropy.apply(y, noise, *params)


def distortion(recon, x):
    """returns (clipped, sse_total, sse[B,2]); sse[:,0] is the UNclamped and sse[:,1] the clamped squared
    error per image, sse_total = sum(sse[:,0]).  Only clipped and sse_total are differentiable."""
    return _Distortion.apply(recon, x)
