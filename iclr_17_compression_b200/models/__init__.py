"""Mirror of the reference's models/__init__.py exports that belong to the Balle-2017 hot path."""
from .GDN import GDN
from .bitEstimator import BitEstimator, Bitparm
from .analysis_17 import Analysis_net_17
from .synthesis_17 import Synthesis_net_17
from .analysis_17_new import Analysis_net_17_new
from .binarizer import Binarizer, Lambda, bin_values
from .ms_ssim_torch import MS_SSIM, SSIM, ms_ssim, ssim

__all__ = ["GDN", "BitEstimator", "Bitparm", "Analysis_net_17", "Synthesis_net_17", "Analysis_net_17_new",
           "Binarizer", "Lambda", "bin_values", "ms_ssim", "ssim", "SSIM", "MS_SSIM"]
