"""Fused training-step tail: the reference clamps every gradient to +-5 element-wise and then runs
torch.optim.Adam (train.py:106-112, 233) - ~120 small launches per step.  FusedAdamClamp does both for all 30
parameter tensors in ONE libb2c launch over flat gradient / moment buffers, and can be replayed inside a CUDA graph
(step count and learning rate live in device memory)."""
import ctypes

import torch

from . import _lib
from .functional import _p, _stream


class FusedAdamClamp(torch.optim.Optimizer):
    """Drop-in for ``optim.Adam(parameters, lr=...)`` followed each step by ``clip_gradient(optimizer, 5)``.

    ``flat_grads``: optional flat fp32 tensor of which every ``p.grad`` is a view, in parameter order (e.g.
    ``GraphedTrainStep.flat``); without it the gradients are gathered into an internal flat buffer first.
    ``param_groups[0]['lr']`` may be changed between steps (train.py:69-81 does)."""

    def __init__(self, params, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, grad_clip=5.0, flat_grads=None):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, grad_clip=grad_clip))
        if len(self.param_groups) != 1:
            raise ValueError("FusedAdamClamp supports a single parameter group")
        ps = self.param_groups[0]["params"]
        if not ps or len(ps) > 32:
            raise ValueError("FusedAdamClamp handles 1..32 parameter tensors")
        for p in ps:
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                raise _lib.B2CError("FusedAdamClamp needs contiguous fp32 CUDA parameters (no CPU path)")
        self._ps = ps
        dev = ps[0].device
        n = sum(p.numel() for p in ps)
        self._flat = flat_grads
        if flat_grads is not None and (flat_grads.numel() != n or flat_grads.dtype != torch.float32):
            raise ValueError("flat_grads must be the fp32 concatenation of all gradients")
        self._own_flat = None
        self.exp_avg = torch.zeros(n, device=dev)
        self.exp_avg_sq = torch.zeros(n, device=dev)
        self.step_t = torch.zeros(1, device=dev, dtype=torch.int32)
        self.lr_t = torch.full((1,), float(lr), device=dev)
        self._lr_host = float(lr)
        self._ptrs = (ctypes.c_void_p * len(ps))(*[p.data_ptr() for p in ps])
        self._sizes = (ctypes.c_size_t * len(ps))(*[p.numel() for p in ps])

    def sync_lr(self):
        """push a changed param_group lr to the device (call outside graph capture)"""
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_host:
            self.lr_t.fill_(lr)
            self._lr_host = lr

    def _grads(self):
        if self._flat is not None:
            return self._flat
        if self._own_flat is None:
            self._own_flat = torch.zeros_like(self.exp_avg)
            o, self._views = 0, []
            for p in self._ps:
                self._views.append(self._own_flat[o:o + p.numel()].view_as(p))
                o += p.numel()
        gs = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self._ps]
        torch._foreach_copy_(self._views, gs)
        return self._own_flat

    @torch.no_grad()
    def step(self, closure=None, capturing=False):
        if closure is not None:
            raise ValueError("closure is not supported")
        if not capturing:
            self.sync_lr()
        g = self.param_groups[0]
        flat = self._grads()
        self.step_t.add_(1)
        lib = _lib.load()
        with torch.cuda.device(flat.device):
            _lib.check(lib.b2c_adam_clamp_step(ctypes.cast(self._ptrs, ctypes.POINTER(ctypes.c_void_p)),
                                               ctypes.cast(self._sizes, ctypes.POINTER(ctypes.c_size_t)), len(self._ps),
                                               _p(flat), _p(self.exp_avg), _p(self.exp_avg_sq), _p(self.step_t),
                                               _p(self.lr_t), float(g["betas"][0]), float(g["betas"][1]),
                                               float(g["eps"]), float(g["grad_clip"]), _stream()), "adam_clamp_step")
        return None
