"""pymf.SNMF on B200: Semi Non-negative Matrix Factorization (pymf/snmf.py:22-90).

The reference class is ``NMF`` with ``update_w`` / ``update_h`` overridden: ``data`` and ``W``
may be signed, ``H`` stays non-negative.  Both overrides consume exactly the reductions of the
NMF passes (``X H^T``, ``H H^T``, ``W^T X``, ``W^T W``), so the engine runs them on the same
kernels with ``pymfb_set_variant(PYMFB_VARIANT_SNMF)`` (include/pymfb.h).
"""
from .nmf import NMF

__all__ = ["SNMF"]


class SNMF(NMF):
    """
    SNMF(data, num_bases=4)

    Semi-NMF: ``|data - W*H|`` minimal with ``H >= 0``.  Drop-in for ``pymf.SNMF``
    (pymf/snmf.py:22): constructor, ``factorize`` and attributes are NMF's; ``update_w`` is
    ``W = X H^T (H H^T)^-1`` (:67-70) and ``update_h`` the square-root ratio of :72-90.
    ``num_bases`` <= 512 (the k x k inverse runs in one CTA, in float64).
    """

    _variant = "snmf"

    @staticmethod
    def _native_hooks():
        return SNMF

    def update_w(self):                                                 # pymf/snmf.py:67-70
        NMF.update_w(self)

    def update_h(self):                                                 # pymf/snmf.py:72-90
        NMF.update_h(self)
