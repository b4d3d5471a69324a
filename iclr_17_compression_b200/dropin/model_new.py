"""Drop-in for the reference's top-level ``model_new`` module (ImageCompressor_new)."""
from iclr_17_compression_b200.model_new import *  # noqa: F401,F403
from iclr_17_compression_b200.model_new import ImageCompressor_new, load_model, save_model  # noqa: F401
