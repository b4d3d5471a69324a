#!/usr/bin/env python
"""Golden vectors for BitEstimator.forward / Bitparm.forward (the factorized CDF as a module call), produced by the
REAL reference classes (/root/reference/models/bitEstimator.py:6-42) on CPU.  Build container only:

    CUDA_VISIBLE_DEVICES="" python tests/golden/make_cdf_golden.py

Stores: CDF values of a seeded [2, N, 5, 7] input under the W1 bit-estimator weights, the output of the first
(non-final) Bitparm layer alone, and the gradients of ``sum(w * CDF(x))`` (seeded w) w.r.t. x and the 11 parameters.
Shim: a stub ``imageio`` module (models/basics.py:15 imports it)."""
import os
import sys
import types

os.environ.setdefault("CUDA_VISIBLE_DEVICES", "")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"

import numpy as np
import torch

sys.modules.setdefault("imageio", types.ModuleType("imageio"))
sys.path.insert(0, REF)
from models.bitEstimator import BitEstimator  # noqa: E402  (the reference's class)

sys.path.insert(0, ROOT)
from oracle import balle17_oracle as O  # noqa: E402  (W1 perturbation only)


def cdf_inputs(N, seed=4004):
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(2, N, 5, 7, generator=g) - 0.5) * 16.0
    w = torch.randn(2, N, 5, 7, generator=g)
    return x, w


def main():
    N = 32
    sd = O.perturb_state_dict(O.init_state_dict(N, seed=234))
    be = BitEstimator(N)
    be.load_state_dict({k[len("bitEstimator."):]: v for k, v in sd.items() if k.startswith("bitEstimator.")})
    x, w = cdf_inputs(N)
    x.requires_grad_(True)
    cdf = be(x)
    (cdf * w).sum().backward()
    out = {"N": N, "cdf": cdf.detach().numpy(), "f1": be.f1(x).detach().numpy(), "dx": x.grad.numpy()}
    for k, p in be.named_parameters():
        out["d/" + k] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "cdf_W1_n32.npz"), **out)
    print("wrote cdf_W1_n32.npz", {k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
