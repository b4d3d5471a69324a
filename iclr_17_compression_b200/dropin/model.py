"""Drop-in for the reference's top-level ``model`` module: ``from model import *`` (train.py:3 and every
evaluation script) resolves to the B200 implementation when this directory precedes the reference
checkout on sys.path (see run_reference_script.py / INTEGRATION.md)."""
from iclr_17_compression_b200.model import *  # noqa: F401,F403
from iclr_17_compression_b200.model import ImageCompressor, load_model, save_model  # noqa: F401
