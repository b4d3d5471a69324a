"""Drop-in for ``models.binarizer``."""
from iclr_17_compression_b200.models.binarizer import Binarizer, Lambda, bin_values  # noqa: F401
