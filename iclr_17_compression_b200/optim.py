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
    ``param_groups[0]['lr']`` may be changed between steps (train.py:69-81 does).

    ``state_dict()`` / ``load_state_dict()`` carry the moments and the step count in torch.optim.Adam's own layout
    (``state[i] = {step, exp_avg, exp_avg_sq}`` per parameter), so a run can resume - also from / into a checkpoint of
    ``torch.optim.Adam`` itself.  Without ``flat_grads``, parameters whose ``.grad`` is None are skipped like Adam
    skips them (no decay of their moments, no update); one step counter is shared by all parameters, so a parameter
    that skipped steps uses the global count in its bias correction (torch keeps one count per parameter)."""

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
        """-> (flat gradient buffer, list of (first, last+1) index runs of parameters that have a gradient)"""
        if self._flat is not None:
            return self._flat, [(0, len(self._ps))]
        if self._own_flat is None:
            self._own_flat = torch.zeros_like(self.exp_avg)
            o, self._views = 0, []
            for p in self._ps:
                self._views.append(self._own_flat[o:o + p.numel()].view_as(p))
                o += p.numel()
        have = [i for i, p in enumerate(self._ps) if p.grad is not None]
        if have:
            torch._foreach_copy_([self._views[i] for i in have], [self._ps[i].grad for i in have])
        runs = []
        for i in have:
            if runs and runs[-1][1] == i:
                runs[-1][1] = i + 1
            else:
                runs.append([i, i + 1])
        return self._own_flat, [tuple(r) for r in runs]

    # ---- checkpointing in torch.optim.Adam's layout (train.py saves only the weights; a resumed run needs these)
    def state_dict(self):
        sd = super().state_dict()
        step = self.step_t.detach().cpu().to(torch.float32).reshape(())
        o, state = 0, {}
        for i, p in enumerate(self._ps):
            n = p.numel()
            state[i] = {"step": step.clone(), "exp_avg": self.exp_avg[o:o + n].view_as(p).clone(),
                        "exp_avg_sq": self.exp_avg_sq[o:o + n].view_as(p).clone()}
            o += n
        sd["state"] = state
        return sd

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        state = state_dict.get("state", {})
        super().load_state_dict({"state": {}, "param_groups": state_dict["param_groups"]})
        for g in self.param_groups:               # a torch.optim.Adam checkpoint has no grad_clip entry
            for k, v in self.defaults.items():
                g.setdefault(k, v)
        o, steps = 0, []
        for i, p in enumerate(self._ps):
            n = p.numel()
            st = state.get(i, state.get(str(i)))
            if st is not None:
                self.exp_avg[o:o + n].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
                steps.append(int(float(st["step"])))
            else:
                self.exp_avg[o:o + n].zero_()
                self.exp_avg_sq[o:o + n].zero_()
            o += n
        self.step_t.fill_(max(steps) if steps else 0)
        self._lr_host = None                      # force the next sync_lr() to push param_groups[0]['lr']
        self.sync_lr()

    # ---- a step in tensor ranges (GraphedTrainStep: each range is launched as soon as its gradients are final)
    @torch.no_grad()
    def begin_step(self):
        """advance the shared step counter once, before the first ``step_range`` of an iteration"""
        self.step_t.add_(1)

    @torch.no_grad()
    def step_range(self, first, last):
        """clamp + Adam for parameter tensors [first, last) of the flat gradient buffer, on the current stream; no
        learning-rate sync and no cache invalidation (the graph-replaying caller does both once per replay)"""
        if self._flat is None:
            raise _lib.B2CError("step_range needs flat_grads")
        g = self.param_groups[0]
        lib = _lib.load()
        vp, sz = ctypes.c_void_p, ctypes.sizeof(ctypes.c_void_p)
        off = sum(p.numel() for p in self._ps[:first]) * 4
        with torch.cuda.device(self._flat.device):
            _lib.check(lib.b2c_adam_clamp_step(
                ctypes.cast(ctypes.byref(self._ptrs, first * sz), ctypes.POINTER(vp)),
                ctypes.cast(ctypes.byref(self._sizes, first * ctypes.sizeof(ctypes.c_size_t)),
                            ctypes.POINTER(ctypes.c_size_t)), last - first,
                vp(self._flat.data_ptr() + off), vp(self.exp_avg.data_ptr() + off), vp(self.exp_avg_sq.data_ptr() + off),
                _p(self.step_t), _p(self.lr_t), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
                float(g["grad_clip"]), _stream()), "adam_clamp_step")

    @torch.no_grad()
    def step(self, closure=None, capturing=False, counted=False):
        """``counted``: the caller has already advanced the step counter (``begin_step``) for this iteration."""
        if closure is not None:
            raise ValueError("closure is not supported")
        if not capturing:
            self.sync_lr()
        g = self.param_groups[0]
        flat, runs = self._grads()
        if not runs:
            return None                                                # no parameter has a gradient: Adam does nothing
        if not counted:
            self.step_t.add_(1)
        lib = _lib.load()
        vp, sz = ctypes.c_void_p, ctypes.sizeof(ctypes.c_void_p)
        with torch.cuda.device(flat.device):
            for first, last in runs:             # one launch per run of consecutive parameters that have a gradient
                off = sum(p.numel() for p in self._ps[:first]) * 4
                _lib.check(lib.b2c_adam_clamp_step(
                    ctypes.cast(ctypes.byref(self._ptrs, first * sz), ctypes.POINTER(vp)),
                    ctypes.cast(ctypes.byref(self._sizes, first * ctypes.sizeof(ctypes.c_size_t)),
                                ctypes.POINTER(ctypes.c_size_t)), last - first,
                    vp(flat.data_ptr() + off), vp(self.exp_avg.data_ptr() + off), vp(self.exp_avg_sq.data_ptr() + off),
                    _p(self.step_t), _p(self.lr_t), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
                    float(g["grad_clip"]), _stream()), "adam_clamp_step")
        # the parameters changed underneath their tensors (no _version bump): invalidate packed-weight caches.  When
        # this call is being captured into a graph the replaying caller bumps once per replay (GraphedTrainStep).
        _lib.bump_weights_epoch()
        return None
