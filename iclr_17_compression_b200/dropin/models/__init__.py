"""Drop-in for the reference's ``models`` package (hot-path exports of models/__init__.py)."""
from iclr_17_compression_b200.models import *  # noqa: F401,F403
from iclr_17_compression_b200.models import GDN, Analysis_net_17, BitEstimator, Bitparm, Synthesis_net_17  # noqa: F401
