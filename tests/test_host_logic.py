"""Host-side logic of pymf_b200.NMF (attribute semantics, flags, early-stop bookkeeping,
hook structure) checked on CPU with the oracle-backed engine double, value-for-value against
the reference's outputs in tests/golden/ref_test_3x50.npz (tests/test_pymf.py:84-95)."""
import logging
import os

import numpy as np
import pytest

import pymf_b200
from oracle import cases
from tests._fake_engine import FakeEngine


@pytest.fixture()
def NMF(monkeypatch):
    monkeypatch.setattr(pymf_b200.NMF, "_engine_factory", FakeEngine)
    return pymf_b200.NMF


def test_reference_test_sequence(NMF, golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_test_3x50.npz"))
    A = cases.ref_test_matrix()
    np.random.seed(cases.REF_TEST_INIT_SEED)
    m = NMF(A, num_bases=4)
    assert not hasattr(m, "W") and not hasattr(m, "H")          # lazy init, pymf/nmf.py:173-177
    m.factorize(show_progress=False, niter=20)
    assert m.ferr.shape == (20,)
    np.testing.assert_allclose(m.ferr, g["ferr_20"], rtol=1e-12)
    assert m.ferr[-1] / (A.shape[0] + A.shape[1]) < 0.1          # tests/test_pymf.py:86-88
    np.testing.assert_allclose(m.W, g["W_20"], rtol=1e-11)
    np.testing.assert_allclose(m.H, g["H_20"], rtol=1e-11)
    m.factorize(compute_h=False)                                  # :92
    np.testing.assert_allclose(m.W, g["W_a"], rtol=1e-11)
    np.testing.assert_allclose(m.H, g["H_a"], rtol=1e-11)
    np.testing.assert_allclose(m.ferr, g["ferr_a"], rtol=1e-12)
    m.factorize(compute_w=False)                                  # :93
    np.testing.assert_allclose(m.H, g["H_b"], rtol=1e-11)
    np.testing.assert_allclose(m.ferr, g["ferr_b"], rtol=1e-12)
    m.factorize(compute_err=False)                                # :94  ferr untouched
    np.testing.assert_allclose(m.ferr, g["ferr_c"], rtol=1e-12)
    np.testing.assert_allclose(m.W, g["W_c"], rtol=1e-11)
    m.factorize(niter=20)                                         # :95  warm start
    np.testing.assert_allclose(m.ferr, g["ferr_d"], rtol=1e-12)
    np.testing.assert_allclose(m.W, g["W_d"], rtol=1e-11)
    np.testing.assert_allclose(m.H, g["H_d"], rtol=1e-11)


def test_early_stop_truncates_like_reference(NMF, golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_conv_3x50.npz"))
    A = cases.ref_test_matrix()
    np.random.seed(cases.REF_TEST_INIT_SEED)
    m = NMF(A, num_bases=4)
    m.factorize(niter=100000)
    assert len(m.ferr) == int(g["ferr_len"])                       # entry i dropped, :201
    np.testing.assert_allclose(m.ferr[-8:], g["ferr_tail"], rtol=1e-10)
    np.testing.assert_allclose(m.W, g["W"], rtol=1e-8)


def test_in_place_identity_and_dtype(NMF):
    X = np.random.RandomState(0).random_sample((6, 9))
    W = np.random.RandomState(1).random_sample((6, 3))
    H = np.random.RandomState(2).random_sample((3, 9))
    W0 = W.copy()
    m = NMF(X, num_bases=3)
    m.W, m.H = W, H
    m.factorize(niter=3)
    assert m.W is W and m.H is H                                   # in-place update, SURVEY 3.4
    assert not np.allclose(W, W0)
    assert m.W.dtype == np.float64
    # list input is accepted and converted
    m2 = NMF(X, num_bases=3)
    m2.W = W0.tolist()
    m2.factorize(niter=1, compute_w=False)
    np.testing.assert_array_equal(m2.W, W0)                         # W frozen


def test_user_mutation_between_calls_is_seen(NMF):
    X = np.random.RandomState(0).random_sample((5, 8))
    np.random.seed(3)
    m = NMF(X, num_bases=2)
    m.factorize(niter=2)
    Hview = m.H
    Hview[0, :] = 0.0                                               # mutate the handed-out array
    m.factorize(niter=1)
    assert np.all(m.H[0, :] == 0.0)                                 # zero rows stay zero under MU


def test_uploads_are_lazy(NMF):
    X = np.random.RandomState(0).random_sample((5, 8))
    np.random.seed(3)
    m = NMF(X, num_bases=2)
    for _ in range(4):
        m.factorize(niter=1)                                        # single-stepping never re-uploads
    assert m._engine.uploads == {"x": 1, "w": 1, "h": 1}
    _ = m.W                                                         # reading hands the array out
    m.factorize(niter=1)
    assert m._engine.uploads == {"x": 1, "w": 2, "h": 1}


def test_sentinel_and_converged(NMF):
    import scipy.sparse
    X = np.random.RandomState(0).random_sample((5, 8))
    m = NMF(X, num_bases=2)
    assert m.frobenius_norm() == -123456                            # pymf/nmf.py:109-112
    ms = NMF(scipy.sparse.csc_matrix(X), num_bases=2)
    ms.W, ms.H = np.ones((5, 2)), np.ones((2, 8))
    assert ms.frobenius_norm() == -123456
    with pytest.raises(TypeError):
        ms.factorize(niter=1)
    m.ferr = np.array([3.0, 2.0, 2.0 + 1e-9])
    assert m.converged(2) and not m.converged(1)


def test_niter_zero_and_signature(NMF):
    X = np.random.RandomState(0).random_sample((5, 8))
    np.random.seed(0)
    m = NMF(X, 2)                                                   # positional num_bases
    m.factorize(0)
    assert m.ferr.shape == (0,) and m.W.shape == (5, 2) and m.H.shape == (2, 8)
    m.factorize(2, False, True, True, True)                         # positional order of the reference
    assert m.ferr.shape == (2,)
    with pytest.raises(TypeError):
        NMF(X, num_bases=2, niter=10)                               # the docstring API is wrong, SURVEY 3.4


def test_overridden_hooks_use_template_loop(NMF):
    calls = []

    class Sub(NMF):
        def update_w(self):
            calls.append("w")
            super(Sub, self).update_w()

    X = np.random.RandomState(0).random_sample((5, 8))
    np.random.seed(1)
    a = Sub(X, num_bases=2)
    a.factorize(niter=3)
    np.random.seed(1)
    b = NMF(X, num_bases=2)
    b.factorize(niter=3)
    assert calls == ["w"] * 3
    np.testing.assert_allclose(a.ferr, b.ferr, rtol=1e-12)
    np.testing.assert_allclose(a.W, b.W, rtol=1e-12)


def test_progress_logging(NMF, caplog):
    X = np.random.RandomState(0).random_sample((5, 8))
    np.random.seed(1)
    m = NMF(X, num_bases=2)
    with caplog.at_level(logging.INFO, logger="pymf"):
        m.factorize(niter=2, show_progress=True)
        assert logging.getLogger("pymf").level == logging.INFO          # pymf/nmf.py:166-169
    msgs = [r.getMessage() for r in caplog.records if r.name == "pymf"]
    assert msgs[0].startswith("Iteration 1/2 FN:") and msgs[1].startswith("Iteration 2/2 FN:")
    m.factorize(niter=1)
    assert logging.getLogger("pymf").level == logging.ERROR


class RecordingSource(object):
    """h5py-like stand-in: `.shape`, `.dtype` and 2-D slicing only; records every slice requested."""

    def __init__(self, a):
        self._a = a
        self.shape = a.shape
        self.dtype = a.dtype
        self.reads = []

    def __getitem__(self, key):
        self.reads.append(key)
        return self._a[key]

    def widest_read(self):
        n = self.shape[1]
        return max(len(range(*k[1].indices(n))) for k in self.reads)


def test_h5py_style_data_source_is_read_in_column_panels(NMF, monkeypatch):
    """The reference touches `data` only through `data[:,:]` and `.shape` (pymf/nmf.py:97,110,125,131 - the h5py
    convention of the package, pymf/kmeans.py:71) and pulls the whole matrix into RAM on every call.  Here any object
    offering `.shape` and slicing is read ONCE, in column panels data[:, c0:c1] of bounded size, never as a whole."""
    import pymf_b200.engine as eng_mod
    rng = np.random.RandomState(3)
    d, n = 9, 1000
    X = rng.random_sample((d, n))
    W0, H0 = rng.random_sample((d, 3)), rng.random_sample((3, n))
    # 128-column panels: 9 rows x 8 bytes x 128 columns
    orig = eng_mod.panel_ranges
    monkeypatch.setattr(eng_mod, "panel_ranges", lambda d_, n_, isz, pb=0: orig(d_, n_, isz, 9 * 8 * 128))
    src = RecordingSource(X)
    a = NMF(src, num_bases=3)
    a.W, a.H = W0.copy(), H0.copy()
    a.factorize(niter=4)
    b = NMF(X, num_bases=3)
    b.W, b.H = W0.copy(), H0.copy()
    b.factorize(niter=4)
    np.testing.assert_array_equal(a.ferr, b.ferr)
    np.testing.assert_array_equal(a.W, b.W)
    assert len(src.reads) == 8 and src.widest_read() == 128           # 7 x 128 + 104 columns, no full-matrix read
    assert all(k[0] == slice(None) for k in src.reads)
    cols = sorted((k[1].start, k[1].stop) for k in src.reads)
    assert cols[0][0] == 0 and cols[-1][1] == n and all(cols[i][1] == cols[i + 1][0] for i in range(len(cols) - 1))
    a.factorize(niter=2)
    assert len(src.reads) == 8                                        # resident: not read again


def test_panel_ranges_cover_the_matrix():
    from pymf_b200.engine import panel_ranges
    pw, r = panel_ranges(16384, 1 << 20, 4)                           # cfg3: 64 MiB panels of 1024 columns
    assert pw == 1024 and len(r) == 1024 and r[0] == (0, 1024) and r[-1] == ((1 << 20) - 1024, 1024)
    pw, r = panel_ranges(3, 50, 8)
    assert pw == 50 and r == [(0, 50)]
    pw, r = panel_ranges(1000, 300, 8, 1000 * 8 * 130)
    assert pw == 128 and r == [(0, 128), (128, 128), (256, 44)]
