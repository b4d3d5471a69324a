"""Drop-in for ``models.analysis_17_new`` (model_new.py:16 imports it by module path)."""
from iclr_17_compression_b200.models.analysis_17_new import Analysis_net_17_new  # noqa: F401
