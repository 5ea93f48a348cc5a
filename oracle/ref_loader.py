"""Load the UNMODIFIED reference implementation of the path (pymf/nmf.py) by file
path.  TEST INFRASTRUCTURE ONLY; works only where a reference checkout exists
(this container: /root/reference).  ``import pymf`` itself fails under Python 3
(cvxopt missing, py2 implicit imports), but pymf/nmf.py only needs numpy/scipy and
one injected name (``xrange``, pymf/nmf.py:182).  Nothing is copied or edited.
"""
import importlib.util
import os

_CANDIDATES = [
    os.environ.get("PYMF_REF", ""),
    "/root/reference/pymf/nmf.py",
]


def find_reference():
    for p in _CANDIDATES:
        if p and os.path.isfile(p):
            return p
    return None


def load_reference_nmf():
    """Returns the reference module (with .NMF) or None if no checkout is present."""
    path = find_reference()
    if path is None:
        return None
    spec = importlib.util.spec_from_file_location("_pymf_ref_nmf", path)
    mod = importlib.util.module_from_spec(spec)
    mod.xrange = range                       # the file's only py2-ism
    spec.loader.exec_module(mod)
    return mod
