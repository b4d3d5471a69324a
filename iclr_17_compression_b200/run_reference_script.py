"""Run an UNMODIFIED script of the reference checkout (train.py, NewTests/*.py ...) on the B200 kernels.

    python -m iclr_17_compression_b200.run_reference_script /path/to/reference/train.py --config cfg.json -n run

The script's ``from model import *`` / ``from models import *`` resolve to iclr_17_compression_b200/dropin;
everything else (datasets.py, Meter.py, ...) still comes from the reference checkout.  Under torchrun each
rank sees only its own GPU (so train.py's nn.DataParallel wrapper degenerates to a pass-through,
train.py:228), gets a rank-dependent --seed unless one is given (train.py:39,202 would otherwise give
every rank the same shuffle), and ImageCompressor instances join the process group and all-reduce their
gradients (ddp.GradSync).  IMAGECOMPRESSOR_RETURN=mse selects the (recon, mse, bpp) tuple train.py:97
unpacks.
"""
import os
import runpy
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit(__doc__)
    script, rest = os.path.abspath(argv[0]), argv[1:]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        os.environ.setdefault("CUDA_VISIBLE_DEVICES", os.environ.get("LOCAL_RANK", "0"))
        os.environ["B2C_DDP"] = "1"
        if os.path.basename(script).startswith("train") and "--seed" not in rest:
            rest += ["--seed", str(234 + rank)]
    if os.path.basename(script) == "train.py":
        os.environ.setdefault("IMAGECOMPRESSOR_RETURN", "mse")
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [os.path.join(here, "dropin"), os.path.dirname(here)]
    sys.path.append(os.path.dirname(script))          # the reference's own helpers, after the drop-in
    sys.argv = [script] + rest
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
