"""iclr_17_compression_b200: the Balle-2017 learned-codec hot path of Yuval-H/iclr_17_compression
(ImageCompressor: analysis transform, quantizer, factorized entropy model, synthesis transform) on
hand-written sm_100a CUDA kernels behind a C ABI (include/b2c.h, libb2c.so)."""
from . import _lib  # noqa: F401
from .model import ImageCompressor, load_model, save_model  # noqa: F401
from .model_new import ImageCompressor_new  # noqa: F401
from .model_twoSteps import ImageCompressor_twoSteps  # noqa: F401
from .models import GDN, Analysis_net_17, Analysis_net_17_new, BitEstimator, Bitparm, Synthesis_net_17  # noqa: F401

__all__ = ["ImageCompressor", "ImageCompressor_new", "ImageCompressor_twoSteps", "load_model", "save_model", "GDN",
           "Analysis_net_17", "Analysis_net_17_new", "Synthesis_net_17", "BitEstimator", "Bitparm"]
