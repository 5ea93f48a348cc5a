"""pymf.BNMF on B200: binary matrix factorization (pymf/bnmf.py:22-123).

The reference class is ``NMF`` with ``update_w`` / ``update_h`` overridden by penalised
ratios and a ``factorize`` that initialises the penalty weights; here the same kernels run
with their ratio epilogue switched by ``pymfb_set_penalty`` (include/pymfb.h), so the whole
loop is still one C call and one pass structure.
"""
from .nmf import NMF

__all__ = ["BNMF"]


class BNMF(NMF):
    """
    BNMF(data, num_bases=4)

    Binary Matrix Factorization: ``|data - W*H|`` minimal with W and H driven towards binary
    values.  Drop-in for ``pymf.BNMF`` (pymf/bnmf.py:22): same constructor, ``factorize``
    signature (note its positional order differs from NMF's, :91-92), hooks and attributes
    (``_lamb_W`` / ``_lamb_H`` / ``_LAMB_INCREASE_W`` / ``_LAMB_INCREASE_H``).
    """

    _LAMB_INCREASE_W = 1.1                                              # pymf/bnmf.py:75
    _LAMB_INCREASE_H = 1.1                                              # pymf/bnmf.py:76

    def _push_penalty(self, eng):
        eng.set_penalty(self._lamb_W, self._lamb_H, self._LAMB_INCREASE_W, self._LAMB_INCREASE_H)

    def _pull_penalty(self, eng):
        self._lamb_W, self._lamb_H = eng.get_penalty()

    def _sync_to_device(self):
        eng = NMF._sync_to_device(self)
        if hasattr(self, "_lamb_W"):
            self._push_penalty(eng)
        return eng

    def update_h(self):                                                 # pymf/bnmf.py:78-85
        NMF.update_h(self)
        self._pull_penalty(self._engine)      # both weights grew by their factor on the device side

    def update_w(self):                                                 # pymf/bnmf.py:86-89
        NMF.update_w(self)

    @staticmethod
    def _native_hooks():
        return BNMF

    def factorize(self, niter=10, compute_w=True, compute_h=True,
                  show_progress=False, compute_err=True):
        """Factorize s.t. WH = data (pymf/bnmf.py:91-123); arguments as in the reference."""
        self._lamb_W = 1.0 / niter                                      # :117
        self._lamb_H = 1.0 / niter                                      # :118
        NMF.factorize(self, niter=niter, compute_w=compute_w, compute_h=compute_h,
                      show_progress=show_progress, compute_err=compute_err)
        if self._engine is not None:
            self._pull_penalty(self._engine)
