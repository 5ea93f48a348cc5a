"""Load the UNMODIFIED reference implementation of the path (pymf/nmf.py) by file
path.  TEST INFRASTRUCTURE ONLY; works only where a reference checkout exists
(this container: /root/reference).  ``import pymf`` itself fails under Python 3
(cvxopt missing, py2 implicit imports), but pymf/nmf.py only needs numpy/scipy and
one injected name (``xrange``, pymf/nmf.py:182).  Nothing is copied or edited.
"""
import importlib.util
import os

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# baseline/_ref is the offline `pip install --target` of the unmodified reference (git-ignored, it travels
# to the GPU box with the repo snapshot; __graft_entry__.build() creates it where /root/reference exists)
_CANDIDATES = [
    os.environ.get("PYMF_REF", ""),
    "/root/reference/pymf/nmf.py",
    os.path.join(_ROOT, "baseline", "_ref", "pymf", "nmf.py"),
]


def find_reference():
    for p in _CANDIDATES:
        if p and os.path.isfile(p):
            return p
    import glob
    for p in sorted(glob.glob(os.path.join(_ROOT, "baseline", "_ref", "**", "pymf", "nmf.py"), recursive=True)):
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


def load_reference_submodule(name):
    """Returns the reference's pymf/<name>.py module (bnmf, snmf) or None.  These files do
    ``from .nmf import NMF``, so they are loaded as submodules of a synthetic, empty package whose
    ``nmf`` member is the by-path module above - pymf/__init__.py (which needs cvxopt) is never
    executed, and nothing is copied or edited."""
    import sys
    import types
    pkg_name = "_pymf_ref_pkg"
    if pkg_name + ".nmf" not in sys.modules:
        nmf = load_reference_nmf()
        if nmf is None:
            return None
        pkg = types.ModuleType(pkg_name)
        pkg.__path__ = [os.path.dirname(find_reference())]
        sys.modules[pkg_name] = pkg
        sys.modules[pkg_name + ".nmf"] = nmf
    full = pkg_name + "." + name
    if full in sys.modules:
        return sys.modules[full]
    pkg_dir = sys.modules[pkg_name].__path__[0]
    spec = importlib.util.spec_from_file_location(full, os.path.join(pkg_dir, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[full] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference_bnmf():
    return load_reference_submodule("bnmf")


def load_reference_snmf():
    return load_reference_submodule("snmf")


def load_reference_nndsvd():
    """pymf/nndsvd.py (and pymf/svd.py, which it imports) from the unmodified reference."""
    load_reference_submodule("svd")
    return load_reference_submodule("nndsvd")
