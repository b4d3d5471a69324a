"""Host-facing inference helper: images in pinned host memory -> latents + per-image statistics back on
the host, with the host<->device copies of one chunk overlapped with the kernels of the previous one
(two CUDA streams, events; no extra threads).  This is the call bench.py times as ``e2e``.

``compress`` takes the reference's fp32 [B,3,H,W] tensors; ``compress_u8`` takes the 8-bit images they were made from
(uint8 [B,3,H,W]) and performs ToTensor's ``/ 255`` on the device - same results bit for bit, a quarter of the PCIe bytes."""
import ctypes

import torch

from . import _lib


class HostPipeline:
    """compress(x_host) for a fixed batch shape; buffers are allocated once."""

    def __init__(self, model, batch, height, width, chunk=8, device=None, use_graphs=True, ramp=True, sizes=None):
        """chunk: images per copy/compute slice.  ramp: the first and the last slice are half-size, so the first
        kernels start (and the last ones finish) sooner next to the PCIe copy that bounds the call.  sizes: an explicit
        slice plan (list of slice sizes summing to batch, each <= chunk) instead."""
        if sizes is not None:
            sizes = [int(v) for v in sizes]
            if sum(sizes) != batch or min(sizes) < 1 or max(sizes) > chunk:
                raise ValueError("sizes must be positive, sum to batch and not exceed chunk")
        elif batch % chunk:
            raise ValueError("batch must be a multiple of chunk")
        else:
            sizes = [chunk] * (batch // chunk)
        if sizes is None or (ramp and chunk >= 2 and len(sizes) >= 2 and len(set(sizes)) == 1 and sizes[0] == chunk):
            h = chunk // 2
            sizes = [h] + [chunk] * (batch // chunk - 1) + [chunk - h]
        self.sizes = sizes
        self.offsets = [sum(sizes[:i]) for i in range(len(sizes))]
        self.model, self.chunk, self.n_chunks = model, chunk, len(sizes)
        self.dev = device or next(model.parameters()).device
        N = model.out_channel_N
        self.lat_host = torch.empty((batch, N, height // 16, width // 16), dtype=torch.float32).pin_memory()
        self.stat_host = torch.empty((batch, 3), dtype=torch.float32).pin_memory()   # bits, sse, sse_clamped
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.out_stream = torch.cuda.Stream(device=self.dev)
        self.out_ready = torch.cuda.Event()
        self.x_dev = [torch.empty((chunk, 3, height, width), device=self.dev) for _ in range(2)]
        self.u8_dev = None                        # allocated by the first compress_u8 call
        self.graphs_u8 = None
        self.h2d_done = [torch.cuda.Event() for _ in range(len(sizes))]     # one per slice: copies are enqueued ahead
        self.buf_free = [torch.cuda.Event() for _ in range(2)]
        self.d2h_done = [torch.cuda.Event() for _ in range(2)]
        self.graphs = None
        if use_graphs:
            self._capture()
        self.h2d_bytes = batch * 3 * height * width * 4
        self.d2h_bytes = self.lat_host.numel() * 4 + self.stat_host.numel() * 4

    @torch.no_grad()
    def _capture(self):
        """one CUDA graph per (input buffer, slice size): the ~14 launches (and their TMA descriptors) of a slice
        replay as a single submission, so small slices are not host-bound"""
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            for b in range(2):
                self.x_dev[b].uniform_(0, 1)
                for n in sorted(set(self.sizes)):
                    self.model(self.x_dev[b][:n])              # warm-up: packs weights, sizes the allocator
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        self.graphs = {}
        for i, n in enumerate(self.sizes):
            b = i & 1
            if (b, n) in self.graphs:
                continue
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                _, z, _ = self.model(self.x_dev[b][:n])
                stats = torch.cat([self.model.last["image_bits"][:, None], self.model.last["sse"]], dim=1)
            self.graphs[(b, n)] = (g, z, stats)

    def _to_f32(self, b, n):
        """device-side ToTensor: u8_dev[b][:n] / 255 -> x_dev[b][:n]"""
        src, dst = self.u8_dev[b][:n], self.x_dev[b][:n]
        with torch.cuda.device(self.dev):
            _lib.check(_lib.load().b2c_u8_to_f32(ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(dst.data_ptr()),
                                                 src.numel(), ctypes.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)),
                       "u8_to_f32")

    @torch.no_grad()
    def _capture_u8(self):
        self.u8_dev = [torch.zeros(self.x_dev[0].shape, device=self.dev, dtype=torch.uint8) for _ in range(2)]
        if self.graphs is None:
            return
        torch.cuda.synchronize(self.dev)
        self.graphs_u8 = {}
        for i, n in enumerate(self.sizes):
            b = i & 1
            if (b, n) in self.graphs_u8:
                continue
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._to_f32(b, n)
                _, z, _ = self.model(self.x_dev[b][:n])
                stats = torch.cat([self.model.last["image_bits"][:, None], self.model.last["sse"]], dim=1)
            self.graphs_u8[(b, n)] = (g, z, stats)

    def compress_u8(self, x_host_u8):
        """x_host_u8: pinned uint8 [batch,3,H,W] (pixel values 0..255).  Same results as compress(x.float() / 255)."""
        if x_host_u8.dtype != torch.uint8:
            raise ValueError("compress_u8 expects uint8 images")
        if self.u8_dev is None:
            self._capture_u8()
        return self.compress(x_host_u8, _u8=True)

    @torch.no_grad()
    def compress(self, x_host, _u8=False):
        """x_host: pinned fp32 [batch,3,H,W].  Returns (latents_host, stats_host) after a full sync."""
        main = torch.cuda.current_stream(self.dev)
        in_dev = self.u8_dev if _u8 else self.x_dev
        graphs = self.graphs_u8 if _u8 else self.graphs

        def enqueue_h2d(i):
            b, n, o = i & 1, self.sizes[i], self.offsets[i]
            with torch.cuda.stream(self.copy_stream):
                if i >= 2:
                    self.copy_stream.wait_event(self.buf_free[b])      # kernels of slice i-2 are done with it
                in_dev[b][:n].copy_(x_host[o:o + n], non_blocking=True)
                self.h2d_done[i].record(self.copy_stream)

        # The host->device copy bounds the call, so the copy engine must never wait for the HOST: the copies of slices
        # 0 and 1 are enqueued before anything else, and the copy of slice i+2 right after the kernels of slice i (it
        # waits, on the device, for the buffer those kernels read).  Enqueued one slice at a time in front of its own
        # kernels, the second copy started 195 us after the first one had finished - the host was still launching the
        # first slice's graph (timeline: profiles/trace_e2e.py).
        for i in range(min(2, self.n_chunks)):
            enqueue_h2d(i)
        for i in range(self.n_chunks):
            b, n, o = i & 1, self.sizes[i], self.offsets[i]
            main.wait_event(self.h2d_done[i])
            if graphs is not None:
                if i >= 2:
                    main.wait_event(self.d2h_done[b])                  # outputs of slice i-2 have left the device
                g, z, stats = graphs[(b, n)]
                g.replay()
            else:
                if _u8:
                    self._to_f32(b, n)
                _, z, _ = self.model(self.x_dev[b][:n])
                stats = torch.cat([self.model.last["image_bits"][:, None], self.model.last["sse"]], dim=1)
            self.buf_free[b].record(main)
            if i + 2 < self.n_chunks:
                enqueue_h2d(i + 2)
            self.out_ready.record(main)
            with torch.cuda.stream(self.out_stream):               # D2H on its own stream (other PCIe direction)
                self.out_stream.wait_event(self.out_ready)
                self.lat_host[o:o + n].copy_(z, non_blocking=True)
                self.stat_host[o:o + n].copy_(stats, non_blocking=True)
                self.d2h_done[b].record(self.out_stream)
                if self.graphs is None:
                    z.record_stream(self.out_stream)
                    stats.record_stream(self.out_stream)
        self.out_stream.synchronize()
        main.synchronize()
        self.model.check_latent_range()          # |z| > 256 guard of the bf16 synthesis path (free: we just synchronised)
        return self.lat_host, self.stat_host
