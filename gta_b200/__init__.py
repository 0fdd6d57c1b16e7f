"""gta_b200 — B200-native geometric-transform-attention (GTA) hot path.

Host-side mirror of the reference operator interface (source/utils/gta.py,
source/utils/wigner_d.py) over a C-ABI CUDA library (include/gta_b200.h).
"""
from .synth import GtaConfig  # noqa: F401

__version__ = "0.1.0"
