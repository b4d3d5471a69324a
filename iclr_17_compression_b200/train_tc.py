"""bf16 tensor-core training step for ImageCompressor(out_channel_N=128): one autograd Function whose
forward runs the fused tcgen05 layers (conv + bias + GDN, saving the GDN norm plane) and whose backward
runs dgrad convs with the fused (I)GDN backward epilogue, tcgen05 weight gradients (MN-major MMAs), and
the fp32 entropy / distortion backward kernels - all through libb2c's C ABI.

Mirrors autograd of model.py:47-80 as driven by train.py:97-105: quantizer = additive uniform noise,
loss terms = sum of squared error on the UNclamped reconstruction and total bits.  Activations between
layers are NHWC bf16; parameters and their gradients are fp32 in the reference's layouts.
Tolerance target (north star): every gradient tensor within 1e-2 relative L2 of the fp32 reference.
"""
import ctypes
import os

import torch

from . import _lib
from . import tc as T
from .functional import _p, _stream, _timed

_vp = ctypes.c_void_p

PARAM_ORDER = [
    "Encoder.conv1.weight", "Encoder.conv1.bias", "Encoder.gdn1.beta", "Encoder.gdn1.gamma",
    "Encoder.conv2.weight", "Encoder.conv2.bias", "Encoder.gdn2.beta", "Encoder.gdn2.gamma",
    "Encoder.conv3.weight",
    "Decoder.deconv1.weight", "Decoder.deconv1.bias", "Decoder.igdn1.beta", "Decoder.igdn1.gamma",
    "Decoder.deconv2.weight", "Decoder.deconv2.bias", "Decoder.igdn2.beta", "Decoder.igdn2.gamma",
    "Decoder.deconv3.weight", "Decoder.deconv3.bias",
    "bitEstimator.f1.h", "bitEstimator.f1.b", "bitEstimator.f1.a",
    "bitEstimator.f2.h", "bitEstimator.f2.b", "bitEstimator.f2.a",
    "bitEstimator.f3.h", "bitEstimator.f3.b", "bitEstimator.f3.a",
    "bitEstimator.f4.h", "bitEstimator.f4.b",
]


def _xsq(y, n, inverse):
    out = torch.empty_like(y)
    with torch.cuda.device(y.device), _timed("tc_xsq", 0, 6 * y.numel()):
        _lib.check(_lib.load().b2c_tc_xsq(_p(y), _p(n), _p(out), y.numel(), int(inverse), _stream()), "tc_xsq")
    return out


class _Pool:
    """Zero-initialised fp32 scratch carved out of ONE flat tensor, so that a whole training step needs a single
    memset instead of one fill kernel per accumulator (weight-gradient partials, column sums, Bitparm gradients)."""

    def __init__(self, flat, offset):
        self.flat, self.off = flat, offset

    def zeros(self, *shape):
        n = 1
        for d in shape:
            n *= d
        if self.off + n > self.flat.numel():
            raise _lib.B2CError("training scratch pool exhausted")
        t = self.flat[self.off:self.off + n].view(*shape)
        self.off += (n + 63) // 64 * 64                     # keep 256-byte alignment (vector red.add)
        return t


def _zeros(pool, dev, *shape):
    return pool.zeros(*shape) if pool is not None else torch.zeros(shape, device=dev, dtype=torch.float32)


def _colsum(t, out=None):
    """out[c] += sum over pixels of an NHWC bf16 tensor; ``out`` must be zero-filled"""
    if out is None:
        out = torch.zeros((128,), device=t.device, dtype=torch.float32)
    with torch.cuda.device(t.device), _timed("tc_nhwc_colsum", 0, 2 * t.numel()):
        _lib.check(_lib.load().b2c_tc_nhwc_colsum(_p(t), _p(out), t.numel() // 128, _stream()), "tc_nhwc_colsum")
    return out


def _gdn_param_grads(beta_p, gamma_p, y, n, dn, inverse, pool=None, out=None, xsq=None, fork=None):
    """d beta_p, d gamma_p from dL/dnorm: d gamma_eff = sum_pix dn_i x_j^2 (tcgen05 wgrad, 1 tap), d beta_eff = sum dn,
    then the LowerBound / square reparametrisation (models/GDN.py:18-24,73-79).  ``xsq``: (x^2 tensor, event) computed
    ahead of time - it depends on the saved forward tensors only.  ``fork``: the column sum of dn runs as a branch of
    its own beside the 1-tap weight gradient (both only read dn)."""
    dev = y.device
    cur = torch.cuda.current_stream(dev)
    if xsq is None:
        xsq = _xsq(y, n, inverse)
    else:
        xsq, ev = xsq
        cur.wait_event(ev)
        xsq.record_stream(cur)
    dbeta_acc = _zeros(pool, dev, 128)
    if fork is not None:
        side = fork.next_stream(avoid=cur)
        side.wait_stream(cur)
        dn.record_stream(side)
        with torch.cuda.stream(side):
            dbeta_eff = _colsum(dn, dbeta_acc)
    dgamma_eff = T.wgrad(dn, xsq, 1, 1, 0, tmp=_zeros(pool, dev, 1, 1, 128, 128)).view(128, 128)
    if fork is not None:
        cur.wait_stream(side)
    else:
        dbeta_eff = _colsum(dn, dbeta_acc)
    dbeta_p, dgamma_p = out if out is not None else (torch.empty_like(beta_p), torch.empty_like(gamma_p))
    with torch.cuda.device(dev):
        _lib.check(_lib.load().b2c_gdn_reparam_bwd(_p(beta_p), _p(gamma_p), _p(dbeta_eff), _p(dgamma_eff), _p(dbeta_p),
                                                   _p(dgamma_p), 128, _stream()), "gdn_reparam_bwd")
    return dbeta_p, dgamma_p


class _Fork:
    """Runs parameter-gradient kernels (weight / bias / GDN gradients, entropy backward) on side streams while the
    dgrad chain continues on the main stream; at batch 8 of 256x256 no single kernel fills 148 SMs, so the
    independent branches are overlapped (also inside a captured CUDA graph, where they become parallel branches)."""
    _streams = {}

    def __init__(self, dev, n=None):
        # One stream per branch (a step has 24 + nested ones): with 3 streams shared round-robin, unrelated branches
        # queued behind each other and the dgrad chain - which joins the entropy backward's stream before the
        # deconv1 dgrad - ended up waiting for a 37 us weight-gradient kernel that happened to share that stream
        # (timeline: profiles/trace_train.py).  The last `n_hi` streams are high priority, for the branches that form
        # the tail of the step (conv1 / gdn1 gradients: nothing else is left to hide them behind).
        if n is None:
            n = int(os.environ.get("B2C_TRAIN_FORKS", "32"))
        self.main = torch.cuda.current_stream(dev)
        key = (dev.index, n)
        if key not in _Fork._streams:
            _Fork._streams[key] = [torch.cuda.Stream(device=dev) for _ in range(n)]
            _Fork._streams[(dev.index, n, "hi")] = [torch.cuda.Stream(device=dev, priority=-1) for _ in range(4)]
        self.sides = _Fork._streams[key]
        self.hi = _Fork._streams[(dev.index, n, "hi")]
        self.i = 0
        self.ih = 0
        self.used = []
        self.outs = []

    @staticmethod
    def _side(dev):
        """a dedicated side stream for operand packing"""
        key = (dev.index, "pack")
        if key not in _Fork._streams:
            _Fork._streams[key] = [torch.cuda.Stream(device=dev)]
        return _Fork._streams[key][0]

    def run(self, fn, *inputs, critical=False):
        if critical and os.environ.get("B2C_TRAIN_TAIL_PRIORITY", "0") == "1":   # measured slower: 308 vs 303 us
            s = self.hi[self.ih % len(self.hi)]
            self.ih += 1
        else:
            s = self.sides[self.i % len(self.sides)]
            self.i += 1
        s.wait_stream(self.main)                  # inputs were produced on the main stream
        for t in inputs:
            t.record_stream(s)
        with torch.cuda.stream(s):
            out = fn()
        self.last_stream = s
        self.last_event = torch.cuda.Event()      # completion of THIS branch (a stream may carry later branches too)
        self.last_event.record(s)
        if s not in self.used:
            self.used.append(s)
        for o in (out if isinstance(out, (tuple, list)) else (out,)):
            self.outs.append(o)
        return out

    def next_stream(self, avoid=None):
        """a side stream for a nested branch (not ``avoid``, the stream the caller is running on); a branch nested in
        a high-priority one is high priority too"""
        if avoid is not None and avoid in self.hi:
            s = self.hi[self.ih % len(self.hi)]
            self.ih += 1
            if s is avoid:
                s = self.hi[self.ih % len(self.hi)]
                self.ih += 1
        else:
            s = self.sides[self.i % len(self.sides)]
            self.i += 1
            if s is avoid:
                s = self.sides[self.i % len(self.sides)]
                self.i += 1
        if s not in self.used:
            self.used.append(s)
        return s

    def join(self):
        for s in self.used:
            self.main.wait_stream(s)
        for o in self.outs:
            o.record_stream(self.main)


def _ksplit():
    """split-K factor of the two 16-tile 5x5 convs of a batch-8 step (conv3 forward, deconv1 dgrad): one CTA per
    output tile would leave 132 of 148 SMs idle, so the 25 taps are spread over KSPLIT CTAs per tile"""
    return max(1, min(25, int(os.environ.get("B2C_TRAIN_KSPLIT", "5"))))


def step_forward(x, noise, prm, pool=None, grad_scale=None, dbias3=None):
    """Forward of the training step on tensor cores (model.py:47-78 with the noise quantizer).  prm: name -> fp32
    CUDA tensor for the 30 parameters.  Returns the state dict the backward needs plus the per-image statistics."""
    lib = _lib.load()
    B, _, H, W = x.shape
    dev = x.device
    W2, W3 = prm["Encoder.conv2.weight"], prm["Encoder.conv3.weight"]
    W1 = prm["Encoder.conv1.weight"]
    D1, D2, D3 = prm["Decoder.deconv1.weight"], prm["Decoder.deconv2.weight"], prm["Decoder.deconv3.weight"]
    # ---- operand packing: only conv1 / gdn1 are needed immediately; every other pack of the step (forward and
    # dgrad forms, gamma and gamma^T) only depends on the parameters and runs on a side stream
    main = torch.cuda.current_stream(dev)
    # operand packing: THREE launches (b2c_tc_pack_many) instead of 21 - conv1's operands on the main stream, the other
    # forward operands and the backward forms on a side stream (as separate launches the forward packs finished 19 us
    # after conv1 did: the second layer waited for them)
    side = _Fork._side(dev)
    side.wait_stream(main)
    (w1p, gp1) = T.pack_many([("w", T.TC_CONV1, W1), ("gdn", prm["Encoder.gdn1.beta"], prm["Encoder.gdn1.gamma"])])
    with torch.cuda.stream(side):
        f = T.pack_many([
            ("w", T.TC_CONV5S2, W2), ("gdn", prm["Encoder.gdn2.beta"], prm["Encoder.gdn2.gamma"]),
            ("w", T.TC_CONV5S2, W3),
            ("w", T.TC_DECONV5S2, D1), ("gdn", prm["Decoder.igdn1.beta"], prm["Decoder.igdn1.gamma"]),
            ("w", T.TC_DECONV5S2, D2), ("gdn", prm["Decoder.igdn2.beta"], prm["Decoder.igdn2.gamma"]),
            ("w", T.deconv3_kind(128), D3)])
        pk = dict(zip(("w2", "gp2", "w3", "d1", "gi1", "d2", "gi2", "d3"), f))
        ev_fwd = torch.cuda.Event()
        ev_fwd.record(side)
        b = T.pack_many([     # backward forms
            ("w", T.TC_CONV1, D3), ("gdn_t", prm["Decoder.igdn2.beta"], prm["Decoder.igdn2.gamma"]),
            ("w", T.TC_CONV5S2, D2), ("gdn_t", prm["Decoder.igdn1.beta"], prm["Decoder.igdn1.gamma"]),
            ("w", T.TC_CONV5S2, D1),
            ("w", T.TC_DECONV5S2, W3), ("gdn_t", prm["Encoder.gdn2.beta"], prm["Encoder.gdn2.gamma"]),
            ("w", T.TC_DECONV5S2, W2), ("gdn_t", prm["Encoder.gdn1.beta"], prm["Encoder.gdn1.gamma"])])
        pk.update(dict(zip(("d3_dg", "gi2_t", "d2_dg", "gi1_t", "d1_dg", "w3_dg", "gp2_t", "w2_dg", "gp1_t"), b)))
        ev_bwd = torch.cuda.Event()
        ev_bwd.record(side)
    for v in pk.values():
        for t in v:
            if t is not None:
                t.record_stream(main)
    # ---- analysis transform
    h1, n1 = T.layer_run(T.TC_CONV1, x, None, w1p, prm["Encoder.conv1.bias"], gdn=gp1, gdn_mode=1,
                         second_plane=True)
    main.wait_event(ev_fwd)
    h2, n2 = T.layer_run(T.TC_CONV5S2, h1, None, pk["w2"], prm["Encoder.conv2.bias"], gdn=pk["gp2"], gdn_mode=1,
                         second_plane=True)
    ks = _ksplit() if h2.shape[0] * h2.shape[1] * h2.shape[2] <= 148 * 128 * 4 else 1      # <= 148 output tiles
    y = T.layer_run(T.TC_CONV5S2, h2, None, pk["w3"], None, out_f32_nchw=True, ksplit=ks,
                    out=(pool.zeros(B, 128, H // 16, W // 16) if (pool is not None and ks > 1) else None))
    # ---- quantizer (noise) + bits
    bitp = [prm[k].reshape(-1) for k in PARAM_ORDER[19:]]
    C, HW = 128, y.shape[2] * y.shape[3]
    z = torch.empty_like(y)
    plane_bits = torch.empty((B, C), device=dev, dtype=torch.float32)
    image_bits = torch.empty((B,), device=dev, dtype=torch.float32)
    pp, keep = _lib.ptr_array(bitp)
    with torch.cuda.device(dev), _timed("quant_entropy_fwd[train]", 60 * y.numel(), 12 * y.numel()):
        _lib.check(lib.b2c_quant_entropy_fwd(_p(y), _p(noise), pp, _p(z), _p(plane_bits), _p(image_bits), B, C, HW,
                                             _stream()), "quant_entropy_fwd")
    # ---- synthesis transform
    zh, _ = T.nchw_to_nhwc(z, False)
    i1, m1 = T.layer_run(T.TC_DECONV5S2, zh, None, pk["d1"], prm["Decoder.deconv1.bias"], gdn=pk["gi1"], gdn_mode=2,
                         second_plane=True)
    i2, m2 = T.layer_run(T.TC_DECONV5S2, i1, None, pk["d2"], prm["Decoder.deconv2.bias"], gdn=pk["gi2"], gdn_mode=2,
                         second_plane=True)
    if grad_scale is not None:
        # distortion forward AND backward fused into the last layer's epilogue: clamp, both squared-error sums and
        # d loss / d recon = grad_scale * (recon - x) leave the kernel together (grad_scale = 2 lambda / numel)
        clipped, sse_partial, g_recon = T.layer_run(T.deconv3_kind(128), i2, None, pk["d3"], prm["Decoder.deconv3.bias"],
                                                    out_f32_nchw=True, ref_image=x, grad_scale=grad_scale,
                                                    reduce_sse=False, dbias=dbias3)
        recon, sse = None, None
    else:
        recon = T.layer_run(T.deconv3_kind(128), i2, None, pk["d3"], prm["Decoder.deconv3.bias"], out_f32_nchw=True)
        # ---- distortion
        clipped = torch.empty_like(recon)
        sse = torch.empty((B, 2), device=dev, dtype=torch.float32)
        nbytes = lib.b2c_clamp_sse_workspace_bytes(B, x[0].numel())
        ws = torch.empty((nbytes,), device=dev, dtype=torch.uint8)
        with torch.cuda.device(dev), _timed("clamp_sse_fwd[train]", 8 * x.numel(), 12 * x.numel()):
            _lib.check(lib.b2c_clamp_sse_fwd(_p(recon), _p(x), _p(clipped), _p(sse), B, x[0].numel(), _p(ws), nbytes,
                                             _stream()), "clamp_sse_fwd")
        sse_partial, g_recon = None, None
    main.wait_event(ev_bwd)          # the backward packs are done long before; keeps stream bookkeeping simple
    return {"prm": prm, "pk": pk, "saved": (x, h1, n1, h2, n2, z, zh, i1, m1, i2, m2, recon),
            "image_bits": image_bits, "plane_bits": plane_bits, "sse": sse, "clipped": clipped, "z": z,
            "sse_partial": sse_partial, "g_recon": g_recon,
            "dbias3_done": dbias3 is not None and grad_scale is not None and T.deconv3_kind(128) == T.TC_DECONV3X}


def step_backward(st, g_sse, g_bits, out=None, pool=None, on_decoder_done=None, g_clipped=None, g_z=None,
                  on_backward_start=None, on_encoder_tail_done=None):
    """Backward of the training step: gradients of ``g_sse * sum (recon - x)^2 + g_bits * total_bits`` for the 30
    parameters (name -> tensor).  ``g_clipped`` / ``g_z``: optional cotangents of the clamped reconstruction and of
    the noisy latent - the route a caller takes who builds the loss from the HEAD return tuple (model.py:80:
    ``lam * mse(clipped, x) + bpp``); clamp passes its gradient where 0 <= recon <= 1, and z = y + noise feeds both
    the synthesis transform and the entropy model, so ``g_z`` simply adds to d loss / d z.  ``out`` (name -> zero-filled fp32 tensor of the parameter's shape, e.g. views of one
    flat buffer) receives the gradients in place; ``pool`` supplies zero-filled scratch.  ``on_decoder_done`` is
    called (on the main stream, with the fork object) once every Decoder.* and bitEstimator.* gradient kernel has
    been enqueued - the hook a data-parallel caller uses to start reducing that bucket while the encoder backward runs.
    ``on_encoder_tail_done(fork)``: the same once the gradients of Encoder.conv2 / gdn2 / conv3 are enqueued (only
    conv1 / gdn1 are still to come); ``on_backward_start()`` runs before the first backward kernel."""
    lib = _lib.load()
    prm, pk = st["prm"], st["pk"]
    x, h1, n1, h2, n2, z, zh, i1, m1, i2, m2, recon = st["saved"]
    dev = x.device
    B = x.shape[0]
    grads = {}

    def dst(k):
        return out[k] if out is not None else None

    def gdn_dst(prefix):
        return (out[prefix + ".beta"], out[prefix + ".gamma"]) if out is not None else None

    fork = _Fork(dev)
    from .functional import _channel_sum_raw
    if on_backward_start is not None:
        on_backward_start()
    # x^2 of the four GDN inputs (operands of the d gamma products) depend on saved forward tensors only: computed now,
    # on side streams, instead of at the head of each parameter-gradient chain (the GDN1 chain is the tail of the step)
    xsqs = {}
    if os.environ.get("B2C_TRAIN_XSQ_EARLY", "1") == "1":
        for key, yy, nn, inv in (("g1", h1, n1, False), ("g2", h2, n2, False), ("i1", i1, m1, True), ("i2", i2, m2, True)):
            t = fork.run(lambda yy=yy, nn=nn, inv=inv: _xsq(yy, nn, inv), yy, nn)
            xsqs[key] = (t, fork.last_event)
    # ---- entropy model: d bits / dz and the 11 parameter vectors (independent of the synthesis chain)
    bitp = [prm[k].reshape(-1) for k in PARAM_ORDER[19:]]
    C, HW = 128, z.shape[2] * z.shape[3]

    def entropy_bwd():
        dbit = [out[k].view(-1) if out is not None else torch.zeros_like(p) for k, p in zip(PARAM_ORDER[19:], bitp)]
        dz_ = torch.empty_like(z)
        pp, k1 = _lib.ptr_array(bitp)
        dp, k2 = _lib.ptr_array(dbit)
        with torch.cuda.device(dev):
            _lib.check(lib.b2c_entropy_bwd(_p(z), pp, _p(g_bits), _p(dz_), dp, B, C, HW, _stream()), "entropy_bwd")
        return [dz_] + dbit
    ent = fork.run(entropy_bwd, g_bits)
    dz, dbit = ent[0], ent[1:]
    ent_done = fork.last_event
    for k, g in zip(PARAM_ORDER[19:], dbit):
        grads[k] = g.view(1, -1, 1, 1)
    # ---- distortion
    g_recon = st.get("g_recon")
    if g_recon is None:
        g_recon = torch.empty_like(recon)
        with torch.cuda.device(dev):
            _lib.check(lib.b2c_clamp_sse_bwd(_p(recon), _p(x), _p(g_clipped), _p(g_sse), _p(g_recon), recon.numel(),
                                             _stream()), "clamp_sse_bwd")
    elif g_clipped is not None:
        raise _lib.B2CError("the fused distortion epilogue (grad_scale) has no cotangent input for the clamped "
                            "reconstruction; run step_forward without grad_scale")
    # ---- deconv3 (9x9/4, 128 -> 3): weight grad (im2col MN-major tcgen05), dgrad (= conv 9x9/4) + IGDN2 backward fused
    grads["Decoder.deconv3.weight"] = fork.run(lambda: T.wgrad9(i2, g_recon, out=dst("Decoder.deconv3.weight")), g_recon)
    if st.get("dbias3_done"):          # accumulated by the forward launch of the last layer (fused distortion backward)
        grads["Decoder.deconv3.bias"] = dst("Decoder.deconv3.bias")
    else:
        grads["Decoder.deconv3.bias"] = fork.run(lambda: _channel_sum_raw(g_recon, dst("Decoder.deconv3.bias")), g_recon)
    gp_d2, dn2 = T.layer_run(T.TC_CONV1, g_recon, None, pk["d3_dg"], None, gdn=pk["gi2_t"],
                             gdn_mode=4, second_plane=True, saved=(i2, m2))
    grads["Decoder.igdn2.beta"], grads["Decoder.igdn2.gamma"] = fork.run(lambda: _gdn_param_grads(
        prm["Decoder.igdn2.beta"], prm["Decoder.igdn2.gamma"], i2, m2, dn2, True, pool, gdn_dst("Decoder.igdn2"),
        xsqs.get("i2"), fork), dn2)
    # ---- deconv2
    grads["Decoder.deconv2.weight"] = fork.run(lambda: T.wgrad(
        i1, gp_d2, 5, 2, 2, tmp=_zeros(pool, dev, 5, 5, 128, 128), out=dst("Decoder.deconv2.weight")), gp_d2)
    grads["Decoder.deconv2.bias"] = fork.run(lambda: _colsum(gp_d2, dst("Decoder.deconv2.bias")), gp_d2)
    gp_d1, dn1 = T.layer_run(T.TC_CONV5S2, gp_d2, None, pk["d2_dg"], None, gdn=pk["gi1_t"],
                             gdn_mode=4, second_plane=True, saved=(i1, m1))
    grads["Decoder.igdn1.beta"], grads["Decoder.igdn1.gamma"] = fork.run(lambda: _gdn_param_grads(
        prm["Decoder.igdn1.beta"], prm["Decoder.igdn1.gamma"], i1, m1, dn1, True, pool, gdn_dst("Decoder.igdn1"),
        xsqs.get("i1"), fork), dn1)
    # ---- deconv1
    grads["Decoder.deconv1.weight"] = fork.run(lambda: T.wgrad(
        zh, gp_d1, 5, 2, 2, tmp=_zeros(pool, dev, 5, 5, 128, 128), out=dst("Decoder.deconv1.weight")), gp_d1)
    grads["Decoder.deconv1.bias"] = fork.run(lambda: _colsum(gp_d1, dst("Decoder.deconv1.bias")), gp_d1)
    if on_decoder_done is not None:
        on_decoder_done(fork)
    fork.main.wait_event(ent_done)                     # dz comes from the entropy backward (first forked task)
    if g_z is not None:
        dz.add_(g_z)                                   # external cotangent of the noisy latent
    ks = _ksplit() if gp_d1.shape[0] * gp_d1.shape[1] * gp_d1.shape[2] <= 148 * 128 * 4 else 1
    if ks > 1:     # split-K CTAs ADD into their output: accumulate on top of d bits / dz (z = y + noise -> dz/dy = 1)
        g_y = T.layer_run(T.TC_CONV5S2, gp_d1, None, pk["d1_dg"], None, out_f32_nchw=True, ksplit=ks, out=dz)
    else:
        g_y = T.layer_run(T.TC_CONV5S2, gp_d1, None, pk["d1_dg"], None, out_f32_nchw=True) + dz
    gyh, _ = T.nchw_to_nhwc(g_y, False)
    # ---- conv3 (no bias): wgrad, dgrad (= transposed conv) + GDN2 backward fused
    grads["Encoder.conv3.weight"] = fork.run(lambda: T.wgrad(
        gyh, h2, 5, 2, 2, tmp=_zeros(pool, dev, 5, 5, 128, 128), out=dst("Encoder.conv3.weight")), gyh)
    gp_c2, dng2 = T.layer_run(T.TC_DECONV5S2, gyh, None, pk["w3_dg"], None, gdn=pk["gp2_t"],
                              gdn_mode=3, second_plane=True, saved=(h2, n2))
    grads["Encoder.gdn2.beta"], grads["Encoder.gdn2.gamma"] = fork.run(lambda: _gdn_param_grads(
        prm["Encoder.gdn2.beta"], prm["Encoder.gdn2.gamma"], h2, n2, dng2, False, pool, gdn_dst("Encoder.gdn2"),
        xsqs.get("g2"), fork), dng2)
    # ---- conv2
    grads["Encoder.conv2.weight"] = fork.run(lambda: T.wgrad(
        gp_c2, h1, 5, 2, 2, tmp=_zeros(pool, dev, 5, 5, 128, 128), out=dst("Encoder.conv2.weight")), gp_c2)
    grads["Encoder.conv2.bias"] = fork.run(lambda: _colsum(gp_c2, dst("Encoder.conv2.bias")), gp_c2)
    if on_encoder_tail_done is not None:
        on_encoder_tail_done(fork)
    gp_c1, dng1 = T.layer_run(T.TC_DECONV5S2, gp_c2, None, pk["w2_dg"], None, gdn=pk["gp1_t"],
                              gdn_mode=3, second_plane=True, saved=(h1, n1))
    grads["Encoder.gdn1.beta"], grads["Encoder.gdn1.gamma"] = fork.run(lambda: _gdn_param_grads(
        prm["Encoder.gdn1.beta"], prm["Encoder.gdn1.gamma"], h1, n1, dng1, False, pool, gdn_dst("Encoder.gdn1"),
        xsqs.get("g1"), fork), dng1, critical=True)
    # ---- conv1 (9x9/4, 3 -> 128): weight grad (im2col MN-major tcgen05); the image needs no gradient
    grads["Encoder.conv1.weight"] = fork.run(lambda: T.wgrad9(gp_c1, x, out=dst("Encoder.conv1.weight")), gp_c1,
                                             critical=True)
    grads["Encoder.conv1.bias"] = fork.run(lambda: _colsum(gp_c1, dst("Encoder.conv1.bias")), gp_c1, critical=True)
    fork.join()
    return grads


class TrainStepTC(torch.autograd.Function):
    """(x, noise, *30 params in PARAM_ORDER) -> (clipped, sse_total, total_bits, z)"""

    @staticmethod
    def forward(ctx, x, noise, *params):
        prm = dict(zip(PARAM_ORDER, [p.detach().contiguous().float() for p in params]))
        st = step_forward(x.detach().contiguous().float(), noise.detach().contiguous().float(), prm)
        total_bits = st["image_bits"].sum()
        sse_total = st["sse"][:, 0].sum()
        ctx.st = st
        ctx.stats = (st["image_bits"], st["plane_bits"], st["sse"])
        # all four outputs are differentiable (model.py:80 hands clipped and z to the caller); unused ones arrive
        # as None in backward instead of materialised zero tensors
        ctx.set_materialize_grads(False)
        return st["clipped"], sse_total, total_bits, st["z"]

    @staticmethod
    def backward(ctx, g_clipped, g_sse, g_bits, g_z):
        st = ctx.st
        if st is None:
            raise _lib.B2CError("TrainStepTC: backward called twice (the saved activations were released)")
        dev = st["z"].device
        zero = torch.zeros((1,), device=dev, dtype=torch.float32)
        g_sse = g_sse.reshape(1).float().contiguous() if g_sse is not None else zero
        g_bits = g_bits.reshape(1).float().contiguous() if g_bits is not None else zero
        g_clipped = g_clipped.float().contiguous() if g_clipped is not None else None
        g_z = g_z.float().contiguous() if g_z is not None else None
        grads = step_backward(st, g_sse, g_bits, g_clipped=g_clipped, g_z=g_z)
        prm = st["prm"]
        ctx.st = None
        return (None, None) + tuple(grads[k].reshape(prm[k].shape) for k in PARAM_ORDER)


class _StepGraphs:
    """Forward and backward of the tensor-core training step captured as TWO CUDA graphs for one input shape and one
    set of parameter tensors, for callers that drive the step through autograd (the reference's unmodified loop:
    ``net(input)`` ... ``rd_loss.backward()``, train.py:97-105).  Eager, that loop is host bound: ~60 libb2c launches
    per step with their TMA descriptors and tensor allocations cost 3.4 ms against 0.36 ms of device time.  The graphs
    read the LIVE parameter tensors (in-place optimizer updates are seen), take the input, the noise and the two loss
    cotangents through static buffers, and keep the saved activations in their private pool."""

    def __init__(self, x, noise, params):
        dev = x.device
        self.x, self.noise = torch.empty_like(x), torch.empty_like(noise)
        self.g_sse = torch.zeros((1,), device=dev, dtype=torch.float32)
        self.g_bits = torch.zeros((1,), device=dev, dtype=torch.float32)
        self.key = _graph_key(x, params)
        prm = dict(zip(PARAM_ORDER, [p.detach() for p in params]))
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):                       # warm-up on a side stream, as graph capture requires
            self.x.copy_(x)
            self.noise.copy_(noise)
            for _ in range(2):
                st = step_forward(self.x, self.noise, prm)
                step_backward(st, self.g_sse, self.g_bits)
            del st
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        self.fwd = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.fwd):
            self.st = step_forward(self.x, self.noise, prm)
            self.total_bits = self.st["image_bits"].sum()
            self.sse_total = self.st["sse"][:, 0].sum()
        self.bwd = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.bwd, pool=self.fwd.pool()):
            grads = step_backward(self.st, self.g_sse, self.g_bits)
            self.flat = torch.cat([grads[k].reshape(-1) for k in PARAM_ORDER])     # one buffer to hand out per step
        self.sizes = [prm[k].numel() for k in PARAM_ORDER]
        self.generation = 0


def _graph_key(x, params):
    return (tuple(x.shape), x.device.index) + tuple(p.data_ptr() for p in params)


class TrainStepTCGraphed(torch.autograd.Function):
    """TrainStepTC as two graph replays: (graphs, x, noise, *30 params) -> (clipped, sse_total, total_bits, z)"""

    @staticmethod
    def forward(ctx, graphs, x, noise, *params):
        graphs.x.copy_(x)
        graphs.noise.copy_(noise)
        graphs.fwd.replay()
        graphs.generation += 1
        ctx.graphs, ctx.generation = graphs, graphs.generation
        ctx.shapes = [p.shape for p in params]
        ctx.set_materialize_grads(False)
        st = graphs.st
        # the graph's buffers are overwritten by the next replay: the caller gets its own copies
        return st["clipped"].clone(), graphs.sse_total.clone(), graphs.total_bits.clone(), st["z"].clone()

    @staticmethod
    def backward(ctx, g_clipped, g_sse, g_bits, g_z):
        G = ctx.graphs
        if G.generation != ctx.generation:
            raise _lib.B2CError("the graphed training step keeps ONE set of saved activations per input shape: "
                                "another training forward of this model ran before this backward.  Call backward "
                                "first, or set model.graph_training = False")
        dev = G.x.device
        zero = torch.zeros((1,), device=dev, dtype=torch.float32)
        g_sse = g_sse.reshape(1).float() if g_sse is not None else zero
        g_bits = g_bits.reshape(1).float() if g_bits is not None else zero
        if g_clipped is not None or g_z is not None:
            # a loss built from the clamped reconstruction / the latent (model.py:80 tuple): eager backward over the
            # activations the forward graph saved
            grads = step_backward(G.st, g_sse.contiguous(), g_bits.contiguous(),
                                  g_clipped=g_clipped.float().contiguous() if g_clipped is not None else None,
                                  g_z=g_z.float().contiguous() if g_z is not None else None)
            return (None, None, None) + tuple(grads[k].reshape(sh) for k, sh in zip(PARAM_ORDER, ctx.shapes))
        G.g_sse.copy_(g_sse)
        G.g_bits.copy_(g_bits)
        G.bwd.replay()
        flat = G.flat.clone()
        outs = [t.view(sh) for t, sh in zip(torch.split(flat, G.sizes), ctx.shapes)]
        return (None, None, None) + tuple(outs)


def _auto_graphs(model, x, noise, params):
    """the _StepGraphs of (shape, parameter tensors), captured the SECOND time that combination is seen (a replica that
    nn.DataParallel rebuilds every step never repeats and stays eager); None = run the eager Function"""
    if not getattr(model, "graph_training", False) or torch.cuda.is_current_stream_capturing():
        return None
    if not all(p.is_contiguous() and p.dtype == torch.float32 for p in params) or x.dtype != torch.float32:
        return None
    cache = model.__dict__.setdefault("_train_graphs", {})
    key = _graph_key(x, params)
    hit = cache.get(key)
    if isinstance(hit, _StepGraphs):
        return hit
    if hit is None:
        if len(cache) >= 4:
            cache.pop(next(iter(cache)))
        cache[key] = 1
        return None
    if hit == "failed":
        return None
    try:
        with torch.no_grad():
            cache[key] = _StepGraphs(x.detach(), noise.detach(), params)
    except Exception as e:                       # capture is an optimisation: the eager step stays correct
        import warnings
        warnings.warn(f"CUDA-graph capture of the training step failed ({e!r}); running it eagerly")
        cache[key] = "failed"
        return None
    return cache[key]


def train_forward(model, x, noise):
    """ImageCompressor training forward on tensor cores: returns (clipped, mse_loss, bpp, z) and fills model.last"""
    params = dict(model.canonical_named_parameters() if hasattr(model, "canonical_named_parameters")
                  else model.named_parameters())
    plist = [params[k] for k in PARAM_ORDER]
    graphs = _auto_graphs(model, x, noise, plist)
    model.__dict__["_last_train_graphed"] = graphs is not None
    if graphs is not None:
        clipped, sse_total, total_bits, z = TrainStepTCGraphed.apply(graphs, x, noise, *plist)
    else:
        clipped, sse_total, total_bits, z = TrainStepTC.apply(x, noise, *plist)
    mse = sse_total / x.numel()
    bpp = total_bits / (x.size(0) * x.size(2) * x.size(3))
    return clipped, mse, bpp, z


class GraphedTrainStep:
    """Whole-step CUDA graph (forward + backward) for a fixed input shape: the kernel launches and their TMA
    descriptors replay as one submission, so the step is not host bound.  The gradients of all 30 tensors live in
    ONE flat fp32 buffer (``p.grad`` are views of it).

    ``direct=True`` (default) drives ``step_forward`` / ``step_backward`` without autograd: the loss weights
    d loss / d sse = lam / numel and d loss / d bits = 1 / (B H W) are constants (train.py:102), every gradient
    kernel writes straight into its slice of the flat buffer, and all accumulators of the step are cleared by a
    single memset.  With ``world_size > 1`` the gradient all-reduce (mean) is captured INSIDE the graph in two
    buckets on a communication stream: Decoder + bitEstimator gradients are reduced over NVLink while the encoder
    backward still runs, the encoder bucket at the end (``allreduce_in_graph=False``: one NCCL call after replay).

        step = GraphedTrainStep(model, x_example, lam)      # after torch.distributed init when multi-GPU
        loss = step(x)                                       # p.grad holds the (averaged) gradients
    """

    def __init__(self, model, x_example, lam=8192.0, warmup=3, group=None, direct=True, allreduce_in_graph=None,
                 optimizer=None, allreduce=None):
        import torch.distributed as dist
        self.model, self.lam, self.group = model, lam, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.x = x_example.clone()
        N = model.out_channel_N
        dev = self.x.device
        B, _, H, W = self.x.shape
        self.noise = torch.zeros(B, N, H // 16, W // 16, device=dev)
        named = (model.canonical_named_parameters() if hasattr(model, "canonical_named_parameters")
                 else list(model.named_parameters()))
        self.params = [p for _, p in named]
        n_par = sum(p.numel() for p in self.params)
        n_par_pad = (n_par + 63) // 64 * 64
        scratch = (4 * 25 * 128 * 128 + 4 * (128 * 128 + 128) + B * N * (H // 16) * (W // 16) + 1024) if direct else 0
        # allreduce (or B2C_AR=...): "nccl" = one NCCL all-reduce after the replay.  "symm" / "symm_graph" /
        # "symm_graph2": the gradient buffer lives in symmetric memory and the mean over ranks is libb2c's own NVLink /
        # NVLS kernel (ddp.SymmAllReduce.all_reduce_fused: cross-rank barrier, reduction and barrier in ONE launch)
        #   symm         launched right after the replay (default when world > 1 and symmetric memory is available)
        #   symm_graph   captured as the LAST node of the step graph (the fused optimizer can then follow inside the
        #                graph: one replay = one whole data-parallel iteration)
        #   symm_graph2  two buckets inside the graph, Decoder + bitEstimator gradients reduced by a few CTAs while the
        #                encoder backward still runs
        # Measured step, batch 8 per GPU (profiles/symm_graph_check.py): 2 GPUs nccl 0.405 / symm 0.387 / symm_graph
        # 0.450 / symm_graph2 0.454 ms; 8 GPUs nccl 0.444 / symm 0.404 / symm_graph2 0.485 ms (1 GPU: 0.357 ms).  Both
        # in-graph forms are SLOWER than the same kernel launched after the replay - a kernel that spins on its peers
        # inside the graph delays the replay by ~60 us - so they stay opt-in.
        self._symm = None
        ar_mode = allreduce or os.environ.get("B2C_AR", "auto")
        if ar_mode == "auto":
            ar_mode = "symm" if (self.world > 1 and direct) else "nccl"
        self._symm_in_graph = False
        self._symm_buckets = 1
        if self.world > 1 and direct and ar_mode in ("symm", "symm_graph", "symm_graph2"):
            from .ddp import SymmAllReduce
            try:
                self._symm = SymmAllReduce(n_par_pad, dev, group, extra=scratch + 64)
            except Exception as e:        # no symmetric memory / peer access on this system: NCCL carries the gradients
                if allreduce is not None or os.environ.get("B2C_AR"):
                    raise
                import warnings
                warnings.warn(f"symmetric-memory all-reduce unavailable ({e!r}); using NCCL")
                self._symm = None
            if self._symm is not None:
                self._symm_in_graph = ar_mode != "symm"
                self._symm_buckets = 2 if ar_mode == "symm_graph2" else 1
        if self._symm is not None:
            # gradients and the step's scratch are ONE contiguous piece of the symmetric allocation: one memset per step
            # (as two tensors the second fill started 9 us after the first had finished, profiles/trace_train_ddp.py)
            self.buf = self._symm.region
            self._scratch_off = self._symm.count
        else:
            self.buf = torch.zeros(n_par_pad + scratch, device=dev, dtype=torch.float32)
            self._scratch_off = n_par_pad
        self._grads = self.buf
        self.flat = self._grads[:n_par]
        views, o = [], 0
        for p in self.params:
            views.append(self.flat[o:o + p.numel()].view_as(p))
            o += p.numel()
        self.views = views
        self.direct = direct
        # optimizer="adam": train.py:106-112 (gradient clamp +-5, Adam) as one fused launch over the flat buffers -
        # captured at the end of the step graph on one GPU, launched after the all-reduce otherwise
        self.optimizer = None
        if optimizer is not None:
            from .optim import FusedAdamClamp
            if optimizer == "adam":
                optimizer = FusedAdamClamp(self.params, flat_grads=self.flat)
            elif not isinstance(optimizer, FusedAdamClamp) or optimizer._flat is not None:
                raise ValueError("optimizer must be 'adam' or a FusedAdamClamp created without flat_grads")
            optimizer._flat = self.flat
            self.optimizer = optimizer
        if allreduce_in_graph is None:
            allreduce_in_graph = os.environ.get("B2C_AR_IN_GRAPH", "0") == "1"
        self.ar_in_graph = bool(allreduce_in_graph and direct and self.world > 1 and self._symm is None)
        self._opt_in_graph = self.optimizer is not None and direct and (self.world == 1 or self.ar_in_graph
                                                                        or self._symm_in_graph)
        if direct:
            from .models import Analysis_net_17
            if type(getattr(model, "_enc", None)) is not Analysis_net_17:
                raise _lib.B2CError("the tensor-core training step implements ImageCompressor / ImageCompressor_twoSteps")
            if N != 128 or set(k for k, _ in named) != set(PARAM_ORDER):
                raise _lib.B2CError("the tensor-core training step needs ImageCompressor(out_channel_N=128)")
            self._capture_direct(named, n_par_pad, warmup)
        else:
            self._capture_autograd(warmup)
        for p, v in zip(self.params, self.views):
            p.grad = v                                   # optimizers read the flat buffer through these views

    # -------------------------------------------------------------------------------------------------------------
    def _run_direct(self, prm, gviews, n_par_pad, enc_elems, with_optimizer=False):
        import torch.distributed as dist
        dev = self.x.device
        main = torch.cuda.current_stream(dev)
        self.buf.zero_()                                              # ONE memset: gradients + all accumulators
        pool = _Pool(self.buf, self._scratch_off)
        fused = os.environ.get("B2C_TRAIN_FUSED_DISTORTION", "1") == "1"
        st = step_forward(self.x, self.noise, prm, pool, grad_scale=(2.0 * self.lam / self.x.numel()) if fused else None,
                          dbias3=gviews["Decoder.deconv3.bias"] if fused else None)
        stat = self._stat_stream
        stat.wait_stream(main)
        with torch.cuda.stream(stat):                                 # reporting only: off the critical path
            self.sse = T.reduce_partial_sse(st["sse_partial"]) if st["sse_partial"] is not None else st["sse"]
            self.mse = self.sse[:, 0].sum() / self.x.numel()           # model.py:61
            self.bpp = st["image_bits"].sum() / (self.x.size(0) * self.x.size(2) * self.x.size(3))   # model.py:78
            self.loss = self.lam * self.mse + self.bpp                # train.py:102
        comm = self._comm_stream

        def decoder_done(fork):
            if not (self.ar_in_graph or (self._symm_in_graph and self._symm_buckets == 2)):
                return
            comm.wait_stream(main)
            for s in fork.used:
                comm.wait_stream(s)
            with torch.cuda.stream(comm):
                if self._symm_in_graph:          # few CTAs: runs beside the encoder backward kernels
                    self._symm.all_reduce_fused(enc_elems, self._symm.count - enc_elems, channel=1,
                                                max_blocks=int(os.environ.get("B2C_AR_BLOCKS", "32")))
                else:
                    dist.all_reduce(self.flat[enc_elems:], op=dist.ReduceOp.AVG, group=self.group)

        # B2C_TRAIN_ADAM_SPLIT=1 (opt-in): the fused clamp + Adam launch of a one-GPU step split into three tensor
        # ranges, each launched on its own branch as soon as its gradients are final (Decoder + bitEstimator: after the
        # deconv1 gradients; conv2 / gdn2 / conv3: before the last dgrad kernel; conv1 / gdn1 at the end).  Measured
        # SLOWER (312 vs 303 us per step): the end of the step is bound by the SM time of the weight-gradient kernels,
        # not by a dependency chain, and two more launches only add to it.
        split_opt = (with_optimizer and self._opt_in_graph and not self.ar_in_graph and not self._symm_in_graph
                     and self._opt_ranges is not None and os.environ.get("B2C_TRAIN_ADAM_SPLIT", "0") == "1")
        n_head, n_enc, n_all = self._opt_ranges if self._opt_ranges is not None else (0, 0, 0)
        opt_stream = self._opt_stream

        def backward_start():
            if with_optimizer and self._opt_in_graph:   # step counter += 1, off the critical path
                opt_stream.wait_stream(main)
                with torch.cuda.stream(opt_stream):
                    self.optimizer.begin_step()

        def opt_range(fork, first, last):
            opt_stream.wait_stream(main)
            for s in fork.used:
                opt_stream.wait_stream(s)
            with torch.cuda.stream(opt_stream):
                self.optimizer.step_range(first, last)

        def decoder_done_all(fork):
            decoder_done(fork)
            if split_opt:
                opt_range(fork, n_enc, n_all)

        def encoder_tail_done(fork):
            if split_opt:
                opt_range(fork, n_head, n_enc)

        step_backward(st, self._g_sse, self._g_bits, out=gviews, pool=pool, on_decoder_done=decoder_done_all,
                      on_backward_start=backward_start, on_encoder_tail_done=encoder_tail_done)
        if split_opt:
            main.wait_stream(opt_stream)
            self.optimizer.step_range(0, n_head)
        if self._symm_in_graph and self._symm_buckets == 1:
            self._symm.all_reduce_fused()                 # last node of the backward pass, on the main stream
        elif self.ar_in_graph or self._symm_in_graph:
            # both buckets go through the SAME stream: collectives of one communicator must be issued in the same
            # order on every rank, and two unordered graph branches would let each rank's replay pick its own order
            comm.wait_stream(main)
            with torch.cuda.stream(comm):
                if self._symm_in_graph:
                    self._symm.all_reduce_fused(0, enc_elems, channel=0)
                else:
                    dist.all_reduce(self.flat[:enc_elems], op=dist.ReduceOp.AVG, group=self.group)
            main.wait_stream(comm)
        if with_optimizer and self._opt_in_graph and not split_opt:
            # inside the graph: gradients are final here.  The step counter was advanced at the start of backward on a
            # side stream, and the (low-priority) chain of small kernels that forms the reported loss is joined AFTER
            # the optimizer - it used to end 4 us after the last gradient kernel and the optimizer waited for it
            main.wait_stream(self._opt_stream)
            self.optimizer.step(capturing=True, counted=True)
        main.wait_stream(stat)

    @torch.no_grad()
    def _capture_direct(self, named, n_par_pad, warmup):
        dev = self.x.device
        B, _, H, W = self.x.shape
        prm = {k: p.detach() for k, p in named}
        for k, t in prm.items():
            if not (t.is_contiguous() and t.dtype == torch.float32):
                raise _lib.B2CError(f"parameter {k} must be contiguous fp32")
        gviews = {k: v for (k, _), v in zip(named, self.views)}
        enc_elems = sum(p.numel() for k, p in named if k.startswith("Encoder."))
        if [k.split(".")[0] for k, _ in named] != sorted([k.split(".")[0] for k, _ in named], key=["Encoder", "Decoder", "bitEstimator"].index):
            raise _lib.B2CError("unexpected parameter order (Encoder, Decoder, bitEstimator expected)")
        # tensor ranges of the split optimizer step: [conv1, gdn1 | rest of the encoder | decoder + entropy model]
        names = [k for k, _ in named]
        head = [k.startswith(("Encoder.conv1.", "Encoder.gdn1.")) for k in names]
        n_head = sum(head)
        self._opt_ranges = (n_head, sum(k.startswith("Encoder.") for k in names), len(names)) \
            if all(head[:n_head]) and 0 < n_head < len(names) else None
        self._g_sse = torch.full((1,), self.lam / self.x.numel(), device=dev, dtype=torch.float32)
        self._g_bits = torch.full((1,), 1.0 / (B * H * W), device=dev, dtype=torch.float32)
        self._stat_stream = torch.cuda.Stream(device=dev)
        self._comm_stream = torch.cuda.Stream(device=dev)
        self._opt_stream = torch.cuda.Stream(device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._run_direct(prm, gviews, n_par_pad, enc_elems)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        # the dgrad chain is captured from a HIGH-priority stream, the parameter-gradient branches from default-
        # priority side streams: kernel nodes inherit the priority, so the block scheduler serves the critical path
        # first when both have CTAs pending (B2C_TRAIN_PRIORITY=0: everything at default priority)
        prio = -1 if os.environ.get("B2C_TRAIN_PRIORITY", "1") == "1" else 0
        cap = torch.cuda.Stream(device=dev, priority=prio)
        with torch.cuda.graph(self.graph, stream=cap):
            self._run_direct(prm, gviews, n_par_pad, enc_elems, with_optimizer=True)   # (warm-ups must not update)
        self.model.last_path = "tcgen05-bf16-train"

    def _capture_autograd(self, warmup):
        model, lam = self.model, self.lam
        dev = self.x.device
        prev_override, prev_mode = model.noise_override, model.return_mode
        model.noise_override, model.return_mode = self.noise, "mse"
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                for p in self.params:
                    p.grad = None
                _, mse, bpp = model(self.x)
                (lam * mse + bpp).backward()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        for p in self.params:
            p.grad = None
        with torch.cuda.graph(self.graph):
            _, self.mse, self.bpp = model(self.x)
            self.loss = lam * self.mse + self.bpp
            self.loss.backward()
            torch._foreach_copy_(self.views, [p.grad for p in self.params])
        model.noise_override, model.return_mode = prev_override, prev_mode

    def __call__(self, x, noise=None):
        """One step on batch ``x``.  ``x`` may be ``self.x`` itself - the graph's static input buffer, e.g. filled by
        the loader's host-to-device copy - in which case nothing is copied (the copy of a resident batch costs 7 us of
        a 290 us step)."""
        if self.optimizer is not None:
            self.optimizer.sync_lr()
        if x is not self.x and x.data_ptr() != self.x.data_ptr():
            self.x.copy_(x, non_blocking=True)
        if noise is None:
            torch.nn.init.uniform_(self.noise, -0.5, 0.5)        # model.py:49
        else:
            self.noise.copy_(noise, non_blocking=True)
        self.graph.replay()
        if self._symm is not None and not self._symm_in_graph:
            self._symm.all_reduce_fused()                            # barrier, NVLink reduce, barrier: one launch
        elif self.world > 1 and not self.ar_in_graph:
            import torch.distributed as dist
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
        if self.optimizer is not None:
            if self._opt_in_graph:
                _lib.bump_weights_epoch()                            # the replay updated the parameters in place
            else:
                self.optimizer.step()
        return self.loss
