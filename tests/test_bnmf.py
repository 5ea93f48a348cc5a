"""BNMF (pymf/bnmf.py), the penalised variant of the NMF loop.

CPU: the numpy restatement reproduces the UNMODIFIED reference's outputs (tests/golden/bnmf_*.npz,
made by `python -m oracle.make_golden bnmf`), and the host class keeps the reference's semantics
(checked with the oracle-backed engine double).  GPU: the CUDA path against the same goldens.
"""
import os

import numpy as np
import pytest

import pymf_b200
from oracle import cases, nmf_oracle as O
from oracle.ref_loader import load_reference_bnmf
from tests._fake_engine import FakeEngine

TOL_WH = 1e-4       # BASELINE.json north_star: per-iteration W/H within 1e-4 relative Frobenius
TOL_FERR = 1e-3     # ... and ferr within 1e-3 relative


def rel(a, b):
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(np.asarray(a, dtype=np.float64) - b) / np.linalg.norm(b))


# --------------------------------------------------------------------------- oracle pin (CPU)
@pytest.mark.parametrize("name", sorted(cases.BNMF_CASES))
def test_bnmf_oracle_matches_reference(name, golden_dir):
    c = cases.BNMF_CASES[name]
    g = np.load(os.path.join(golden_dir, "%s.npz" % name))
    X, W, H = cases.build(name)
    X = X.astype(np.float64)
    snaps = {}

    def rec(i, W_, H_, f):
        if (i + 1) in c["keep"]:
            snaps[i + 1] = (W_.copy(), H_.copy())

    lam = {}
    ferr = O.bnmf_factorize(X, W, H, niter=c["niter"], early_stop=False, record=rec, lam=lam)
    np.testing.assert_allclose(ferr, g["ferr"], rtol=1e-12)
    np.testing.assert_allclose([lam["W"], lam["H"]], g["lam_steps"], rtol=1e-15)
    tol = 1e-6 if c.get("store32") else 1e-12
    for it, (W_, H_) in snaps.items():
        assert rel(W_, g["W_%d" % it]) < tol and rel(H_, g["H_%d" % it]) < tol
    # the whole-call fixture (early stop active)
    X, W, H = cases.build(name)
    f = O.bnmf_factorize(X.astype(np.float64), W, H, niter=c["niter"])
    np.testing.assert_allclose(f, g["ferr_whole"], rtol=1e-12)
    assert rel(W, g["Wf"]) < tol and rel(H, g["Hf"]) < tol


def test_bnmf_oracle_against_live_reference_if_present():
    refb = load_reference_bnmf()
    if refb is None:
        pytest.skip("no reference checkout on this box")
    rng = np.random.RandomState(3)
    X = (rng.random_sample((19, 33)) < 0.4).astype(np.float64)
    W0, H0 = rng.random_sample((19, 3)), rng.random_sample((3, 33))
    m = refb.BNMF(X, num_bases=3)
    m.W, m.H = W0.copy(), H0.copy()
    m.factorize(niter=25)
    W, H = W0.copy(), H0.copy()
    lam = {}
    ferr = O.bnmf_factorize(X, W, H, niter=25, lam=lam)
    np.testing.assert_allclose(ferr, m.ferr, rtol=1e-13)
    np.testing.assert_allclose(W, m.W, rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(H, m.H, rtol=1e-12, atol=1e-300)
    assert lam["H"] == m._lamb_H and lam["W"] == m._lamb_W
    # the second call restarts the weights at 1/niter (pymf/bnmf.py:117-118) and H-only keeps W
    m.factorize(niter=5, compute_w=False)
    f2 = O.bnmf_factorize(X, W, H, niter=5, compute_w=False)
    np.testing.assert_allclose(f2, m.ferr, rtol=1e-13)
    np.testing.assert_allclose(H, m.H, rtol=1e-12, atol=1e-300)


# --------------------------------------------------------------------------- host logic (CPU)
@pytest.fixture()
def FakeBNMF(monkeypatch):
    monkeypatch.setattr(pymf_b200.NMF, "_engine_factory", FakeEngine)
    return pymf_b200.BNMF


def test_bnmf_host_class_semantics(FakeBNMF, golden_dir):
    name = "bnmf_bin"
    c = cases.BNMF_CASES[name]
    g = np.load(os.path.join(golden_dir, "%s.npz" % name))
    X, W0, H0 = cases.build(name)
    m = FakeBNMF(X, num_bases=c["k"])
    m.W, m.H = W0.copy(), H0.copy()
    m.factorize(c["niter"])                         # niter is the first positional argument (:91)
    np.testing.assert_allclose(m.ferr, g["ferr_whole"], rtol=1e-12)
    assert rel(m.W, g["Wf"]) < 1e-12 and rel(m.H, g["Hf"]) < 1e-12
    np.testing.assert_allclose([m._lamb_W, m._lamb_H], g["lam_whole"], rtol=1e-15)
    # positional order of BNMF.factorize: (niter, compute_w, compute_h, show_progress, compute_err)
    Wb = m.W.copy()
    m.factorize(3, False, True, False, False)
    np.testing.assert_array_equal(m.W, Wb)
    assert abs(m._lamb_H - (1.0 / 3) * 1.1 ** 3) < 1e-15
    # default niter is 10 (:91), NMF's is 1
    m.factorize(compute_err=False)
    assert abs(m._lamb_H - 0.1 * 1.1 ** 10) < 1e-15


def test_bnmf_hooks_step_like_the_reference_loop(FakeBNMF, golden_dir):
    """Driving update_w / update_h / frobenius_norm by hand (what NMF.factorize does, pymf/nmf.py:182-190)
    gives the stepped fixture, and a subclass overriding a hook falls back to that template loop."""
    name = "bnmf_ragged"
    c = cases.BNMF_CASES[name]
    g = np.load(os.path.join(golden_dir, "%s.npz" % name))
    X, W0, H0 = cases.build(name)
    m = FakeBNMF(X, num_bases=c["k"])
    m.W, m.H = W0.copy(), H0.copy()
    m._lamb_W = m._lamb_H = 1.0 / c["niter"]
    ferr = []
    for _ in range(c["niter"]):
        m.update_w()
        m.update_h()
        ferr.append(m.frobenius_norm())
    np.testing.assert_allclose(ferr, g["ferr"], rtol=1e-12)
    np.testing.assert_allclose([m._lamb_W, m._lamb_H], g["lam_steps"], rtol=1e-15)

    calls = []

    class Traced(FakeBNMF):
        def update_h(self):
            calls.append("h")
            FakeBNMF.update_h(self)

    t = Traced(X, num_bases=c["k"])
    t.W, t.H = W0.copy(), H0.copy()
    t.factorize(niter=c["niter"])
    assert len(calls) == len(t.ferr) + (1 if len(t.ferr) < c["niter"] else 0)
    np.testing.assert_allclose(t.ferr, g["ferr_whole"], rtol=1e-12)


# --------------------------------------------------------------------------- CUDA path (GPU)
def _paths(d, n, k):
    return ["simt", "auto"] if (d >= 64 and n >= 128) else ["simt"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(cases.BNMF_CASES))
def test_bnmf_gpu_trajectory_matches_reference_golden(name, golden_dir):
    c = cases.BNMF_CASES[name]
    g = np.load(os.path.join(golden_dir, "%s.npz" % name))
    X, W0, H0 = cases.build(name)
    for path in _paths(c["d"], c["n"], c["k"]):
        m = pymf_b200.BNMF(X, num_bases=c["k"], path=path)
        m.W, m.H = W0.copy(), H0.copy()
        m._lamb_W = m._lamb_H = 1.0 / c["niter"]
        ferr = np.zeros(c["niter"])
        for i in range(c["niter"]):                 # single-stepped through the hooks
            m.update_w()
            m.update_h()
            ferr[i] = m.frobenius_norm()
            if (i + 1) in c["keep"]:
                assert rel(m.W, g["W_%d" % (i + 1)]) < TOL_WH, (path, i)
                assert rel(m.H, g["H_%d" % (i + 1)]) < TOL_WH, (path, i)
        assert np.max(np.abs(ferr - g["ferr"]) / g["ferr"]) < TOL_FERR
        np.testing.assert_allclose([m._lamb_W, m._lamb_H], g["lam_steps"], rtol=1e-12)
        # whole call = one C call for all iterations
        w = pymf_b200.BNMF(X, num_bases=c["k"], path=path)
        w.W, w.H = W0.copy(), H0.copy()
        w.factorize(niter=c["niter"])
        assert len(w.ferr) == len(g["ferr_whole"])
        assert np.max(np.abs(w.ferr - g["ferr_whole"]) / g["ferr_whole"]) < TOL_FERR
        assert rel(w.W, g["Wf"]) < TOL_WH and rel(w.H, g["Hf"]) < TOL_WH
        np.testing.assert_allclose([w._lamb_W, w._lamb_H], g["lam_whole"], rtol=1e-12)


@pytest.mark.gpu
def test_bnmf_drives_factors_binary_on_planted_data():
    """The property the method exists for (pymf/bnmf.py:70-76): with growing lambda the factors end up
    near {0, 1}; and lambda = 0 through the same entry point is plain NMF (same code path; equal up to the
    summation order of the fp32 atomics in the X.H^T flush, which varies from run to run)."""
    X, W0, H0 = cases.build("bnmf_bin")
    m = pymf_b200.BNMF(X, num_bases=6)
    m.W, m.H = W0.copy(), H0.copy()
    m.factorize(niter=150)                          # the oracle ends fully binary here (lambda ~ 1e4)
    assert np.isfinite(m.W).all() and np.isfinite(m.H).all()
    for F in (m.W, m.H):
        assert np.mean(np.minimum(np.abs(F), np.abs(F - 1.0)) < 0.05) > 0.95
    a = pymf_b200.NMF(X, num_bases=6)
    a.W, a.H = W0.copy(), H0.copy()
    a.factorize(niter=5)
    b = pymf_b200.NMF(X, num_bases=6)
    b.W, b.H = W0.copy(), H0.copy()
    b._sync_to_device().set_penalty(0.0, 0.0, 1.1, 1.1)
    b.factorize(niter=5)
    assert rel(a.W, b.W) < 2e-6 and rel(a.H, b.H) < 2e-6


# --------------------------------------------------------------------------- the reference's own BNMF test vector
def _ref_sequence(cls, g, tol_wh, tol_f):
    """tests/test_pymf.py:80,84-95 for BNMF: np.round(A - 2.0), k = 4, niter = 20, then the flag runs."""
    A = np.round(cases.ref_test_matrix() - 2.0)
    np.random.seed(cases.REF_TEST_INIT_SEED)
    m = cls(A, num_bases=4)
    m.factorize(show_progress=False, niter=20)
    assert m.ferr[-1] / (A.shape[0] + A.shape[1]) < 0.1                 # the reference's bound, :86-88
    assert np.max(np.abs(m.ferr - g["ferr_20"]) / g["ferr_20"]) < tol_f
    assert rel(m.W, g["W_20"]) < tol_wh and rel(m.H, g["H_20"]) < tol_wh
    m.factorize(show_progress=False, compute_h=False, niter=20)         # :92
    assert rel(m.W, g["W_a"]) < tol_wh and np.max(np.abs(m.ferr - g["ferr_a"]) / g["ferr_a"]) < tol_f
    m.factorize(show_progress=False, compute_w=False, niter=20)         # :93
    assert rel(m.H, g["H_b"]) < tol_wh and np.max(np.abs(m.ferr - g["ferr_b"]) / g["ferr_b"]) < tol_f
    m.factorize(show_progress=False, compute_err=False, niter=20)       # :94 (ferr untouched)
    assert rel(m.W, g["W_c"]) < tol_wh and rel(m.H, g["H_c"]) < tol_wh
    assert len(m.ferr) == len(g["ferr_c"])
    m.factorize(show_progress=False, niter=20)                          # :95 warm start
    assert len(m.ferr) == len(g["ferr_d"])
    assert rel(m.W, g["W_d"]) < tol_wh and rel(m.H, g["H_d"]) < tol_wh
    np.testing.assert_allclose([m._lamb_W, m._lamb_H], g["lam_d"], rtol=1e-12)


def test_bnmf_reference_test_sequence_host_logic(FakeBNMF, golden_dir):
    _ref_sequence(FakeBNMF, np.load(os.path.join(golden_dir, "bnmf_ref_test_3x50.npz")), 1e-11, 1e-11)


@pytest.mark.gpu
def test_bnmf_reference_test_sequence_on_gpu(golden_dir):
    _ref_sequence(pymf_b200.BNMF, np.load(os.path.join(golden_dir, "bnmf_ref_test_3x50.npz")), TOL_WH, TOL_FERR)
