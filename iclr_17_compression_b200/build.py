"""Builds libb2c.so (the C-ABI CUDA library, include/b2c.h) in-tree with nvcc for sm_100a.

    python -m iclr_17_compression_b200.build [--force] [--verbose]

The .so is linked against the static CUDA runtime only (no libcuda / torch link), so it
loads on a machine without a driver (symbol checks on the CPU build box) and travels to
the GPU box with the repo snapshot.
"""
import glob
import hashlib
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "libb2c.so")
STAMP = os.path.join(PKG, "build", "libb2c.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
    "--expt-relaxed-constexpr",
    "-cudart", "static",
] + os.environ.get("B2C_EXTRA_NVCC_FLAGS", "").split()      # e.g. -DB2C_TCP_FINE_TRACE (debug timelines)


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _digest():
    h = hashlib.sha256()
    for p in _sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(ROOT, "include", "b2c.h"), __file__]:
        h.update(p.encode())
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current():
    try:
        return os.path.exists(LIB) and open(STAMP).read().strip() == _digest()
    except OSError:
        return False


def build(force=False, verbose=False):
    if not force and is_current():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for src in _sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {os.path.basename(src)} (exit {p.returncode})\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libb2c.so")
    cmd = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
    subprocess.run(cmd, check=True)
    with open(STAMP, "w") as f:
        f.write(_digest())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
