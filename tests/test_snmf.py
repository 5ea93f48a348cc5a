"""SNMF (pymf/snmf.py), the semi-non-negative variant of the NMF loop.

CPU: the numpy restatement reproduces the UNMODIFIED reference's outputs (tests/golden/snmf_*.npz,
made by `python -m oracle.make_golden snmf`) and the host class keeps the reference's semantics
(oracle-backed engine double).  GPU: the CUDA path against the same goldens.
"""
import os

import numpy as np
import pytest

import pymf_b200
from oracle import cases, nmf_oracle as O
from oracle.ref_loader import load_reference_snmf
from tests._fake_engine import FakeEngine

TOL_WH = 1e-4       # BASELINE.json north_star: per-iteration W/H within 1e-4 relative Frobenius
TOL_FERR = 1e-3


def rel(a, b):
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(np.asarray(a, dtype=np.float64) - b) / np.linalg.norm(b))


def ferr_close(a, b, tol):
    return len(a) == len(b) and (len(b) == 0 or np.max(np.abs(np.asarray(a) - b) / b) < tol)


# --------------------------------------------------------------------------- oracle pin (CPU)
@pytest.mark.parametrize("name", sorted(cases.SNMF_CASES))
def test_snmf_oracle_matches_reference(name, golden_dir):
    c = cases.SNMF_CASES[name]
    g = np.load(os.path.join(golden_dir, "%s.npz" % name))
    X, W, H = cases.build(name)
    snaps = {}

    def rec(i, W_, H_, f):
        if (i + 1) in c["keep"]:
            snaps[i + 1] = (W_.copy(), H_.copy())

    W, ferr = O.snmf_factorize(X.astype(np.float64), W, H, niter=c["niter"], early_stop=False, record=rec)
    np.testing.assert_allclose(ferr, g["ferr"], rtol=1e-11)
    tol = 1e-6 if c.get("store32") else 1e-10
    for it, (W_, H_) in snaps.items():
        assert rel(W_, g["W_%d" % it]) < tol and rel(H_, g["H_%d" % it]) < tol


def test_snmf_oracle_against_live_reference_if_present():
    refs = load_reference_snmf()
    if refs is None:
        pytest.skip("no reference checkout on this box")
    rng = np.random.RandomState(9)
    X = rng.standard_normal((17, 45))
    W0, H0 = rng.random_sample((17, 3)), rng.random_sample((3, 45))
    m = refs.SNMF(X, num_bases=3)
    m.W, m.H = W0.copy(), H0.copy()
    m.factorize(niter=25)
    H = H0.copy()
    W, ferr = O.snmf_factorize(X, W0.copy(), H, niter=25)
    np.testing.assert_allclose(ferr, m.ferr, rtol=1e-12)
    np.testing.assert_allclose(W, m.W, rtol=1e-10)
    np.testing.assert_allclose(H, m.H, rtol=1e-10, atol=1e-300)


# --------------------------------------------------------------------------- the reference's own test vector
def _ref_sequence(cls, g, tol_wh, tol_f):
    """tests/test_pymf.py:77,84-95 for SNMF: A = rand(3, 50) + 2, k = 4, niter = 20, then the flag runs
    (the compute_h=False run stops early: W = X H^T (H H^T)^-1 does not move once H is fixed)."""
    A = cases.ref_test_matrix()
    np.random.seed(cases.REF_TEST_INIT_SEED)
    m = cls(A, num_bases=4)
    m.factorize(show_progress=False, niter=20)
    assert m.ferr[-1] / (A.shape[0] + A.shape[1]) < 0.1                 # the reference's bound, :86-88
    assert ferr_close(m.ferr, g["ferr_20"], tol_f)
    assert rel(m.W, g["W_20"]) < tol_wh and rel(m.H, g["H_20"]) < tol_wh
    m.factorize(show_progress=False, compute_h=False, niter=20)         # :92
    assert rel(m.W, g["W_a"]) < tol_wh and ferr_close(m.ferr, g["ferr_a"], tol_f)
    m.factorize(show_progress=False, compute_w=False, niter=20)         # :93
    assert rel(m.H, g["H_b"]) < tol_wh and ferr_close(m.ferr, g["ferr_b"], tol_f)
    m.factorize(show_progress=False, compute_err=False, niter=20)       # :94
    assert rel(m.W, g["W_c"]) < tol_wh and rel(m.H, g["H_c"]) < tol_wh
    m.factorize(show_progress=False, niter=20)                          # :95
    assert ferr_close(m.ferr, g["ferr_d"], tol_f)
    assert rel(m.W, g["W_d"]) < tol_wh and rel(m.H, g["H_d"]) < tol_wh


@pytest.fixture()
def FakeSNMF(monkeypatch):
    monkeypatch.setattr(pymf_b200.NMF, "_engine_factory", FakeEngine)
    return pymf_b200.SNMF


def test_snmf_reference_test_sequence_host_logic(FakeSNMF, golden_dir):
    _ref_sequence(FakeSNMF, np.load(os.path.join(golden_dir, "snmf_ref_test_3x50.npz")), 1e-9, 1e-9)


def test_snmf_subclass_hook_falls_back_to_template_loop(FakeSNMF, golden_dir):
    name = "snmf_signed"
    c = cases.SNMF_CASES[name]
    g = np.load(os.path.join(golden_dir, "%s.npz" % name))
    X, W0, H0 = cases.build(name)
    calls = []

    class Traced(FakeSNMF):
        def update_w(self):
            calls.append("w")
            FakeSNMF.update_w(self)

    t = Traced(X, num_bases=c["k"])
    t.W, t.H = W0.copy(), H0.copy()
    t.factorize(niter=c["niter"])
    assert len(calls) >= len(t.ferr)
    np.testing.assert_allclose(t.ferr, g["ferr"][:len(t.ferr)], rtol=1e-10)


# --------------------------------------------------------------------------- CUDA path (GPU)
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(cases.SNMF_CASES))
def test_snmf_gpu_trajectory_matches_reference_golden(name, golden_dir):
    c = cases.SNMF_CASES[name]
    g = np.load(os.path.join(golden_dir, "%s.npz" % name))
    X, W0, H0 = cases.build(name)
    paths = ["simt", "auto"] if (c["d"] >= 64 and c["n"] >= 128) else ["simt"]
    for path in paths:
        m = pymf_b200.SNMF(X, num_bases=c["k"], path=path)
        m.W, m.H = W0.copy(), H0.copy()
        ferr = np.zeros(c["niter"])
        for i in range(c["niter"]):                 # single-stepped: factorize(niter=1) never stops early
            m.factorize(niter=1)
            ferr[i] = m.ferr[0]
            if (i + 1) in c["keep"]:
                assert rel(m.W, g["W_%d" % (i + 1)]) < TOL_WH, (path, i)
                assert rel(m.H, g["H_%d" % (i + 1)]) < TOL_WH, (path, i)
        assert np.max(np.abs(ferr - g["ferr"]) / g["ferr"]) < TOL_FERR
        assert m.H.min() >= 0.0                     # the "semi" in semi-NMF
        # whole call = one C call
        w = pymf_b200.SNMF(X, num_bases=c["k"], path=path)
        w.W, w.H = W0.copy(), H0.copy()
        w.factorize(niter=c["niter"])
        n = len(w.ferr)
        assert np.max(np.abs(w.ferr - g["ferr"][:n]) / g["ferr"][:n]) < TOL_FERR
        if n == c["niter"]:
            last = max(c["keep"])
            assert rel(w.W, g["W_%d" % last]) < TOL_WH and rel(w.H, g["H_%d" % last]) < TOL_WH


@pytest.mark.gpu
def test_snmf_reference_test_sequence_on_gpu(golden_dir):
    _ref_sequence(pymf_b200.SNMF, np.load(os.path.join(golden_dir, "snmf_ref_test_3x50.npz")), TOL_WH, TOL_FERR)


@pytest.mark.gpu
def test_snmf_wide_k_and_limit():
    """k = 160 (> 128: several 32-wide k blocks, a 160 x 160 fp64 inverse) against the oracle; k > 512 is rejected."""
    rng = np.random.RandomState(0)
    d, n, k = 200, 300, 160
    X = rng.random_sample((d, n)) - 0.3
    W0, H0 = rng.random_sample((d, k)), rng.random_sample((k, n))
    m = pymf_b200.SNMF(X, num_bases=k)
    m.W, m.H = W0.copy(), H0.copy()
    m.factorize(niter=3)
    Hr = H0.copy()
    Wr, fr = O.snmf_factorize(X, W0.copy(), Hr, niter=3)
    # cond(H H^T) ~ 6e3 here: W = A B^-1 amplifies the fp32 round-off of A and B (a float32 numpy simulation of the
    # same steps lands at 3.5e-5 for W and 2e-6 for H), hence the wider bound on W
    assert rel(m.H, Hr) < TOL_WH and rel(m.W, Wr) < 5e-4
    assert np.max(np.abs(m.ferr - fr) / fr) < TOL_FERR
    big = pymf_b200.SNMF(rng.random_sample((600, 700)), num_bases=513)
    with pytest.raises(pymf_b200.PymfbError, match="k <= 512"):
        big.factorize(niter=1)
