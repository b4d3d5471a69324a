"""Helpers shared by the oracle-vs-golden (CPU) and CUDA-vs-oracle (GPU) tests."""
import glob
import os

import numpy as np
import torch

from oracle import balle17_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names(prefix=""):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


def digest(t):
    t = t.detach().double().flatten().cpu()
    idx = torch.linspace(0, t.numel() - 1, steps=min(64, t.numel())).long()
    return np.array([t.sum().item(), t.abs().sum().item(), t.norm().item()]), t[idx].numpy()


_sd_cache = {}


def weights(N, name):
    key = (N, name)
    if key not in _sd_cache:
        sd = O.init_state_dict(N, seed=234)
        if name == "W1":
            sd = O.perturb_state_dict(sd)
        _sd_cache[key] = sd
    return {k: v.clone() for k, v in _sd_cache[key].items()}


class Case:
    def __init__(self, name):
        self.name = name
        self.g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        g = self.g
        self.N, self.B, self.H, self.W = int(g["N"]), int(g["B"]), int(g["H"]), int(g["W"])
        self.training = bool(int(g["training"]))
        self.lam = float(g["lam"])
        self.sd = weights(self.N, str(g["weights"]))
        self.x = O.synth_images(str(g["kind"]), self.B, self.H, self.W, int(g["xseed"]))
        self.noise = O.synth_noise(self.B, self.N, self.H, self.W, int(g["nseed"])) if self.training else None

    def check_weights(self):
        """regenerated weights == the reference's (digest recorded by make_golden.py)"""
        for k, v in self.sd.items():
            s, smp = digest(v)
            np.testing.assert_allclose(s, self.g["w_sum/" + k], rtol=1e-12, atol=0, err_msg=k)
            np.testing.assert_array_equal(smp, self.g["w_smp/" + k], err_msg=k)


class NewCase:
    """fixtures of the 'next' rows: ImageCompressor_new / ImageCompressor_twoSteps, W1 weights"""

    def __init__(self, name):
        self.name = name
        self.g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
        g = self.g
        self.N, self.B, self.H, self.W = int(g["N"]), int(g["B"]), int(g["H"]), int(g["W"])
        self.sd = weights(self.N, "W1")
        self.x = O.synth_images(str(g["kind"]), self.B, self.H, self.W, int(g["xseed"]))
        if "tseed" in g.files:
            gen = torch.Generator().manual_seed(int(g["tseed"]))
            self.target = (torch.rand(self.B, self.N, self.H // 16, self.W // 16, generator=gen) > 0.5).float()
