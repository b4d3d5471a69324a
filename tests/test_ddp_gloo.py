"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: gradient all-reduce buckets, parameter
broadcast, per-image sharding and the host-side statistics gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class Toy(torch.nn.Module):
    """same prefix structure as ImageCompressor so DEFAULT_BUCKETS applies"""

    def __init__(self):
        super().__init__()
        self.Encoder = torch.nn.ModuleDict({"conv1": torch.nn.Linear(4, 4), "conv2": torch.nn.Linear(4, 4),
                                            "conv3": torch.nn.Linear(4, 4, bias=False)})
        self.Decoder = torch.nn.ModuleDict({"deconv1": torch.nn.Linear(4, 4), "deconv3": torch.nn.Linear(4, 2)})
        self.bitEstimator = torch.nn.ModuleDict({"f1": torch.nn.Linear(4, 1)})
        self.unused = torch.nn.Linear(2, 2)              # never gets a gradient

    def forward(self, x, rate=True):
        h = self.Encoder["conv3"](self.Encoder["conv2"](self.Encoder["conv1"](x)))
        d = self.Decoder["deconv3"](self.Decoder["deconv1"](h)).pow(2).mean()
        return d + self.bitEstimator["f1"](h).mean() if rate else d


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from iclr_17_compression_b200.ddp import GradSync, gather_stats, shard_range
        torch.manual_seed(100 + rank)                    # replicas start DIFFERENT
        m = Toy()
        m.unused.requires_grad_(False)
        sync = GradSync(m, auto_finish=(rank >= 0 and os.environ.get("B2C_TEST_AUTO") == "1"))
        ref = Toy()
        ref.load_state_dict(m.state_dict())              # after broadcast: rank 0's weights everywhere
        w0 = [torch.zeros_like(p) for p in m.parameters()]
        for t, p in zip(w0, m.parameters()):
            t.copy_(p.data)
            dist.broadcast(t, src=0)
        same = all(torch.equal(t, p.data) for t, p in zip(w0, m.parameters()))
        xs = [torch.randn(3, 4, generator=torch.Generator().manual_seed(7 + r)) for r in range(world)]
        for step in range(2):                            # two steps: buckets reset correctly
            for p in m.parameters():
                p.grad = None
            m(xs[rank] * (step + 1)).backward()
            if not sync.auto_finish:
                sync.finish()
        # expected: mean over ranks of the local gradients
        exp = None
        for r in range(world):
            ref.zero_grad()
            ref(xs[r] * 2).backward()
            g = [p.grad.clone() if p.grad is not None else None for p in ref.parameters()]
            exp = g if exp is None else [a + b if a is not None else None for a, b in zip(exp, g)]
        ok = True
        for p, e in zip(m.parameters(), exp):
            if e is None:
                ok &= p.grad is None
            else:
                ok &= torch.allclose(p.grad, e / world, rtol=1e-6, atol=1e-7)
        # a step in which some parameters get NO gradient (loss without the rate term) must still be
        # finished and averaged, and must not poison the following steps
        def expected(rate, scale):
            acc = None
            for r in range(world):
                ref.zero_grad(set_to_none=True)
                ref(xs[r] * scale, rate=rate).backward()
                g = [p.grad.clone() if p.grad is not None else None for p in ref.parameters()]
                acc = g if acc is None else [a + b if a is not None else None for a, b in zip(acc, g)]
            return acc
        ok_partial = True
        for rate, scale in ((False, 3.0), (True, 4.0), (False, 5.0)):
            for p in m.parameters():
                p.grad = None
            m(xs[rank] * scale, rate=rate).backward()
            if not sync.auto_finish:
                sync.finish()
            for p, e in zip(m.parameters(), expected(rate, scale)):
                if e is None:
                    ok_partial &= p.grad is None
                else:
                    ok_partial &= torch.allclose(p.grad, e / world, rtol=1e-6, atol=1e-7)
            ok_partial &= all(b["ready"] == 0 and b["work"] is None for b in sync.buckets)
        ok &= ok_partial
        # bucket prefixes are canonical: ImageCompressor_twoSteps (Encoder1 / Decoder1) gets the same three buckets
        from iclr_17_compression_b200.model_twoSteps import ImageCompressor_twoSteps
        m2 = ImageCompressor_twoSteps(8)
        s2 = GradSync(m2, broadcast_parameters=False)
        ok &= len(s2.buckets) == 3 and all(n.startswith("Decoder.") for n, _ in s2.buckets[0]["members"])
        ok &= sum(len(b["members"]) for b in s2.buckets) == 30
        s2.remove()
        s, e = shard_range(24, rank, world)
        stats = gather_stats(torch.arange(s, e, dtype=torch.float32)[:, None])
        q.put((rank, same, ok, len(sync.buckets), (s, e), stats.flatten().tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("auto", ["0", "1"])
def test_gradsync_and_sharding_world2(auto):
    os.environ["B2C_TEST_AUTO"] = auto
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same, ok, nb, rng, stats in sorted(res):
        assert same, "parameters were not broadcast from rank 0"
        assert ok, "all-reduced gradients differ from the mean of the per-rank gradients"
        assert nb == 3                                    # decoder / bits + encoder tail / encoder head
        assert rng == (rank * 12, rank * 12 + 12)
        assert stats == [float(i) for i in range(24)]


def test_shard_range_is_a_partition():
    from iclr_17_compression_b200.ddp import shard_range
    for n in (0, 1, 7, 24, 25):
        for w in (1, 2, 3, 8):
            cuts = [shard_range(n, r, w) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            assert max(b - a for a, b in cuts) - min(b - a for a, b in cuts) <= 1


def test_dropin_modules_resolve():
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path[:0]=[%r,%r];"
            "from model import *; import models;"
            "assert ImageCompressor.__module__.startswith('iclr_17_compression_b200');"
            "assert all(n in globals() for n in ['load_model','save_model','logging','np','nn','F','math','torch']);"
            "assert models.GDN.__module__.startswith('iclr_17_compression_b200'); print('ok')"
            % (os.path.join(root, "iclr_17_compression_b200", "dropin"), root))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr
