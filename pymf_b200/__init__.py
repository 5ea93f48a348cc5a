"""pymf_b200 - B200-native implementation of pymf's NMF multiplicative-update path
(and BNMF / SNMF, the penalised and the semi-non-negative variants of the same loop, and the
NNDSVD initialisation that feeds it).

``pymf_b200.NMF`` mirrors ``pymf.NMF`` (pymf/nmf.py); the compute runs in libpymfb.so
(hand-written sm_100a CUDA behind the C ABI of include/pymfb.h).
"""
from .nmf import NMF  # noqa: F401
from .bnmf import BNMF  # noqa: F401
from .snmf import SNMF  # noqa: F401
from .nndsvd import NNDSVD  # noqa: F401
from .engine import Engine, pinned_empty, numa_info  # noqa: F401
from ._lib import PymfbError  # noqa: F401

__all__ = ["NMF", "BNMF", "SNMF", "NNDSVD", "Engine", "PymfbError", "pinned_empty", "numa_info"]
__version__ = "0.1.0"
