"""Mirror of the reference's models/__init__.py exports that belong to the Balle-2017 hot path."""
from .GDN import GDN
from .bitEstimator import BitEstimator, Bitparm
from .analysis_17 import Analysis_net_17
from .synthesis_17 import Synthesis_net_17

__all__ = ["GDN", "BitEstimator", "Bitparm", "Analysis_net_17", "Synthesis_net_17"]
