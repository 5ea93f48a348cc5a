"""Drop-in alias: ``import pymf; pymf.NMF(...)`` resolves to the B200 implementation.

Only the NMF multiplicative-update path of the reference package (NMF and its BNMF / SNMF
update-rule variants) is provided
(pymf/__init__.py:16-43 re-exports ~20 other factorizations - out of scope).
"""
from pymf_b200 import NMF, BNMF, SNMF  # noqa: F401

__all__ = ["NMF", "BNMF", "SNMF"]
