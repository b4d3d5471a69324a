"""Drop-in for the reference's top-level ``model_twoSteps`` module (ImageCompressor_twoSteps)."""
from iclr_17_compression_b200.model_twoSteps import *  # noqa: F401,F403
from iclr_17_compression_b200.model_twoSteps import ImageCompressor_twoSteps, load_model, save_model  # noqa: F401
