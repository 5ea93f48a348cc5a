"""The C-ABI library loads and exports every symbol include/pymfb.h declares (CPU only:
no compute call is made), and the product fails loudly without a GPU."""
import os
import re

import numpy as np
import pytest

from pymf_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "pymfb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pymfb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "libpymfb.so does not export %s" % s
    # and the ctypes table covers exactly the header
    assert sorted(_lib.SIGNATURES) == syms


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.pymfb_version() >= 1000
    assert isinstance(lib.pymfb_last_error(), bytes)


def test_no_cpu_fallback_without_gpu():
    if _lib.device_count() > 0:
        pytest.skip("a GPU is visible; the no-GPU behaviour is tested on the CPU box")
    import pymf_b200
    with pytest.raises(pymf_b200.PymfbError, match="no CUDA device"):
        pymf_b200.Engine(8, 8, 2)
    m = pymf_b200.NMF(np.random.random((4, 6)), num_bases=2)
    with pytest.raises(pymf_b200.PymfbError):
        m.factorize(niter=1)


def test_product_does_not_import_oracle():
    """The product package must never reach into oracle/ (parity claims depend on it)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pymf_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f


def test_pinned_empty_fails_loudly_without_gpu():
    if _lib.device_count() > 0:
        pytest.skip("a GPU is visible")
    import pymf_b200
    with pytest.raises(pymf_b200.PymfbError, match="no CUDA device"):
        pymf_b200.pinned_empty((4, 4))
