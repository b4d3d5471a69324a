"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed over NCCL/NVLink.

Training: pure data parallel.  The reference wraps the model in single-process
nn.DataParallel (train.py:228: broadcast parameters + reduce gradients to GPU 0 every step); here every
rank keeps a replica and ``GradSync`` all-reduces the 7.07 MB of fp32 gradients in a few buckets in
reverse layer order, each launched asynchronously from a post-accumulate-grad hook the moment its last
gradient is produced, so the transfer overlaps the rest of backward.  ``finish()`` waits, averages and
hands the gradients back before ``optimizer.step()``.

Inference: images are independent -> ``shard_range`` splits them across ranks; there is no data-path
collective, only a host-side gather of per-image statistics (``gather_stats``).

Works with any backend (the CPU tests use gloo, world_size 2).
"""
import os
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """contiguous, balanced [start, stop) of this rank (first n_items % world ranks get one extra)."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_stats(local: torch.Tensor, group=None) -> torch.Tensor:
    """host-side gather of per-image rows [n_local, k] from every rank -> [n_total, k] (all ranks)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local.detach().cpu()
    objs: List[Optional[torch.Tensor]] = [None] * dist.get_world_size(group)
    dist.all_gather_object(objs, local.detach().cpu(), group=group)
    return torch.cat(objs, dim=0)


# reverse layer order = order in which backward produces the gradients (decoder first)
DEFAULT_BUCKETS: Sequence[Sequence[str]] = (
    ("Decoder.",),
    ("bitEstimator.", "Encoder.conv3", "Encoder.gdn2", "Encoder.conv2"),
    ("Encoder.gdn1", "Encoder.conv1"),
)


class GradSync:
    """Bucketed, overlapped gradient all-reduce (mean) for a replicated module."""

    def __init__(self, module: torch.nn.Module, buckets: Optional[Sequence[Sequence[str]]] = None, group=None,
                 broadcast_parameters: bool = True, auto_finish: bool = False):
        if not dist.is_initialized():
            raise RuntimeError("GradSync needs an initialised torch.distributed process group")
        self.group = group
        self.world = dist.get_world_size(group)
        # auto_finish: finish() runs inside the hook that launches the last bucket, so an unmodified training
        # loop (train.py:105-112: backward(); clip; optimizer.step()) sees averaged gradients.
        self.auto_finish = auto_finish
        # bucket prefixes are written against the canonical Encoder. / Decoder. names (ImageCompressor_twoSteps
        # stores the transforms as Encoder1 / Decoder1, model_twoSteps.py:41-42)
        named_all = (module.canonical_named_parameters() if hasattr(module, "canonical_named_parameters")
                     else list(module.named_parameters()))
        named = [(n, p) for n, p in named_all if p.requires_grad]
        prefixes = buckets if buckets is not None else DEFAULT_BUCKETS
        assigned = {}
        self.buckets = []
        for prefs in prefixes:
            members = [(n, p) for n, p in named if n not in assigned and any(n.startswith(pf) for pf in prefs)]
            if members:
                self._add_bucket(members, assigned)
        rest = [(n, p) for n, p in named if n not in assigned]
        if rest:
            self._add_bucket(rest, assigned)
        if broadcast_parameters:                      # replicas start identical (rank 0's weights)
            with torch.no_grad():
                for _, p in module.named_parameters():
                    dist.broadcast(p.data, src=0, group=group)
                for _, b in module.named_buffers():
                    dist.broadcast(b.data, src=0, group=group)
        self._handles = []
        self._armed = False
        for bi, b in enumerate(self.buckets):
            for j, (n, p) in enumerate(b["members"]):
                self._handles.append(p.register_post_accumulate_grad_hook(self._make_hook(bi, j)))

    def _add_bucket(self, members, assigned):
        numel = sum(p.numel() for _, p in members)
        p0 = members[0][1]
        flat = torch.zeros(numel, dtype=p0.dtype, device=p0.device)
        offs, o = [], 0
        for n, p in members:
            offs.append(o)
            o += p.numel()
            assigned[n] = len(self.buckets)
        self.buckets.append({"members": members, "flat": flat, "offsets": offs, "ready": 0, "work": None})

    def _arm_backward_callback(self):
        """run ``_backward_done`` when the autograd pass that fired this hook ends, however many gradients it
        produced: a step in which some parameter gets no gradient (a loss without the rate term leaves
        bitEstimator.* untouched) must still wait for, average and hand back the buckets that were launched, and
        must reset the per-step counters"""
        if self._armed:
            return
        try:
            torch.autograd.Variable._execution_engine.queue_callback(self._backward_done)
            self._armed = True
        except RuntimeError:            # hook fired outside an autograd pass: the caller finishes explicitly
            self._armed = False

    def _backward_done(self):
        self._armed = False
        if self.auto_finish:
            self.finish()

    def _make_hook(self, bi, j):
        def hook(p):
            self._arm_backward_callback()
            b = self.buckets[bi]
            o = b["offsets"][j]
            b["flat"][o:o + p.numel()].copy_(p.grad.reshape(-1))
            b["ready"] += 1
            if b["ready"] == len(b["members"]) and b["work"] is None:
                b["work"] = dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        return hook

    def finish(self):
        """wait for every bucket, average, write the reduced gradients back into .grad; resets the step state.
        Buckets that did not fill up (some member got no gradient) are reduced now with zeros in the gaps.  EVERY
        bucket is reduced in every step that produced any gradient, in bucket order, so all ranks issue the same
        sequence of collectives whatever subset of parameters their graph touched.  A second call in the same step
        (explicit finish() after an auto-finished backward) is a no-op."""
        if all(b["ready"] == 0 and b["work"] is None for b in self.buckets):
            return
        for b in self.buckets:
            if b["work"] is None:
                b["work"] = dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        for b in self.buckets:
            b["work"].wait()
            b["flat"].div_(self.world)
            for (n, p), o in zip(b["members"], b["offsets"]):
                if p.grad is not None:
                    p.grad.copy_(b["flat"][o:o + p.numel()].view_as(p))
            b["ready"] = 0
            b["work"] = None
            b["flat"].zero_()

    def reduce_now(self):
        """no-hook variant (e.g. after a CUDA-graph replay of forward+backward): pack, all-reduce and average
        every bucket from the current .grad tensors"""
        for b in self.buckets:
            for (n, p), o in zip(b["members"], b["offsets"]):
                if p.grad is not None:
                    b["flat"][o:o + p.numel()].copy_(p.grad.reshape(-1))
            b["ready"] = len(b["members"])
            if b["work"] is None:
                b["work"] = dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.finish()

    def bucket_bytes(self):
        return [b["flat"].numel() * b["flat"].element_size() for b in self.buckets]

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []


class SymmAllReduce:
    """In-place mean of a flat fp32 buffer over NVLink / NVSwitch peer memory (libb2c ``b2c_allreduce_symm``).

    The buffer is allocated in torch's symmetric memory (the same allocation is mapped into every rank of the
    group); ``all_reduce()`` enqueues, on the current stream: cross-rank barrier -> one kernel in which this rank
    reduces its 1/world slice (NVLS ``multimem.ld_reduce`` + ``multimem.st`` when the switch supports multicast,
    peer loads / stores otherwise) -> barrier.  For the 7 MB gradient buffer of the training step this replaces a
    ~58 us NCCL all-reduce (protocol latency bound at this size).  CUDA only; there is no fallback."""

    def __init__(self, numel: int, device, group=None, use_multicast: Optional[bool] = None, extra: int = 0):
        """``extra``: fp32 elements of caller scratch allocated right behind the reduced buffer (``self.region`` =
        buffer + scratch as one contiguous tensor, so that a training step clears both with a single memset)."""
        import ctypes
        import torch.distributed._symmetric_memory as symm
        if not dist.is_initialized():
            raise RuntimeError("SymmAllReduce needs an initialised torch.distributed process group")
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        from . import _lib
        q = 4 * self.world
        self.count = (numel + q - 1) // q * q
        # behind the data: flag blocks of the fused kernel (barrier -> reduce -> barrier in one launch), one per channel
        self.flag_words = int(_lib.load().b2c_allreduce_symm_flag_words())
        self.channels = 4
        self.extra = (int(extra) + 63) // 64 * 64
        self.flag_base = self.count + self.extra
        self._alloc = symm.empty(self.flag_base + self.channels * self.flag_words, dtype=torch.float32, device=device)
        self._alloc.zero_()
        self.buf = self._alloc[:self.count]
        self.region = self._alloc[:self.flag_base]
        self.hdl = symm.rendezvous(self._alloc, self.group.group_name)
        torch.cuda.synchronize(device)
        self.hdl.barrier(channel=0)                 # every rank's flags are zero before anybody's first launch
        if use_multicast is None:
            # in-switch reduction (NVLS multimem) from three ranks on; between TWO GPUs plain peer loads / stores are
            # faster (7 MB: 27.9 vs 37.7 us per call, profiles/symm_micro.py).  B2C_AR_MC=0 / 1 overrides.
            env = os.environ.get("B2C_AR_MC")
            use_multicast = (env == "1") if env in ("0", "1") else self.world > 2
        mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0) if use_multicast else 0
        self.multicast = mc != 0
        self._mc = ctypes.c_void_p(mc)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self._ptrs = (ctypes.c_void_p * self.world)(*ptrs)
        self._ctypes = ctypes

    def all_reduce(self, mean: bool = True):
        from . import _lib
        ct = self._ctypes
        self.hdl.barrier(channel=0)
        with torch.cuda.device(self.buf.device):
            _lib.check(_lib.load().b2c_allreduce_symm(ct.cast(self._ptrs, ct.POINTER(ct.c_void_p)), self._mc, self.rank,
                                                      self.world, self.count, ct.c_float(1.0 / self.world if mean else 1.0),
                                                      ct.c_void_p(torch.cuda.current_stream(self.buf.device).cuda_stream)),
                       "allreduce_symm")
        self.hdl.barrier(channel=1)

    def all_reduce_fused(self, begin: int = 0, count: Optional[int] = None, channel: int = 0, mean: bool = True,
                         max_blocks: int = 0):
        """ONE launch on the current stream: cross-rank barrier, reduction of elements [begin, begin + count), cross-rank
        barrier - the flags live behind the data in the symmetric buffer (``channel`` selects the flag block: use one
        per bucket).  Capturable in a CUDA graph.  Every rank must issue the same sequence of calls per channel."""
        from . import _lib
        count = self.count - begin if count is None else count
        if begin % 4 or count % 4 or count <= 0 or begin + count > self.count or not 0 <= channel < self.channels:
            raise ValueError("all_reduce_fused: begin / count must be multiples of 4 inside the buffer, channel < 4")
        ct = self._ctypes
        with torch.cuda.device(self.buf.device):
            _lib.check(_lib.load().b2c_allreduce_symm_fused(
                ct.cast(self._ptrs, ct.POINTER(ct.c_void_p)), self._mc, self.rank, self.world, begin, count,
                self.flag_base + channel * self.flag_words, ct.c_float(1.0 / self.world if mean else 1.0), int(max_blocks),
                ct.c_void_p(torch.cuda.current_stream(self.buf.device).cuda_stream)), "allreduce_symm_fused")

    def all_reduce_range(self, begin: int, count: int, mean: bool = True, max_blocks: int = 0):
        """the same for elements [begin, begin + count) only (multiples of 4), with at most ``max_blocks`` CTAs: a
        bucket of the gradient buffer, reducible while other kernels of the step still run.  Enqueued on the current
        stream: barrier -> reduce -> barrier; every rank must issue its range calls in the same order."""
        from . import _lib
        if begin % 4 or count % 4 or count <= 0 or begin + count > self.count:
            raise ValueError("all_reduce_range: begin / count must be multiples of 4 inside the buffer")
        ct = self._ctypes
        self.hdl.barrier(channel=0)
        with torch.cuda.device(self.buf.device):
            _lib.check(_lib.load().b2c_allreduce_symm_range(
                ct.cast(self._ptrs, ct.POINTER(ct.c_void_p)), self._mc, self.rank, self.world, begin, count,
                ct.c_float(1.0 / self.world if mean else 1.0), int(max_blocks),
                ct.c_void_p(torch.cuda.current_stream(self.buf.device).cuda_stream)), "allreduce_symm_range")
        self.hdl.barrier(channel=1)
