"""NNDSVD initialisation (pymf/nndsvd.py:79-108).  The fixtures tests/golden/nndsvd_*.npz come from the
unmodified reference (oracle/make_golden.py nndsvd).  CPU: the oracle restatement reproduces them, and the closed form
the device uses for the reference's second SVD is the same thing.  GPU: pymf_b200.NNDSVD against the fixtures, and the
reference NMF trajectory warm-started from them."""
import os

import numpy as np
import pytest

import pymf_b200
from oracle import cases, nmf_oracle as O

TOL_WH = 1e-4
TOL_FERR = 1e-3


def rel(a, b):
    b = np.asarray(b, dtype=np.float64)
    return np.linalg.norm(np.asarray(a, dtype=np.float64) - b) / np.linalg.norm(b)


@pytest.mark.parametrize("name", sorted(cases.NNDSVD_CASES))
def test_oracle_nndsvd_matches_reference(name, golden_dir):
    c = cases.NNDSVD_CASES[name]
    g = np.load(os.path.join(golden_dir, "%s.npz" % name))
    X = cases.build_nndsvd(name)
    W, H = O.nndsvd(X, c["k"])
    assert rel(W, g["W"]) < 1e-9 and rel(H, g["H"]) < 1e-9
    assert abs(O.frobenius_norm(X, W, H) - g["ferr"][0]) / g["ferr"][0] < 1e-10


@pytest.mark.parametrize("name", ["nndsvd_small", "nndsvd_tall"])
def test_closed_form_of_the_second_svd(name, golden_dir):
    """max(0, s u v^T) = s (u+ v+^T + u- v-^T) with disjoint supports: its top singular triplet is the larger of
    (s |u+||v+|, u+/|u+|, v+/|v+|) and the same of the negative parts - what kernels_svd.cuh computes instead of the
    reference's second d x n SVD (pymf/nndsvd.py:94-108)."""
    c = cases.NNDSVD_CASES[name]
    g = np.load(os.path.join(golden_dir, "%s.npz" % name))
    X = cases.build_nndsvd(name)
    U, S, V = O.svd_dense(X)
    W = np.zeros_like(g["W"]); H = np.zeros_like(g["H"])
    W[:, 0] = np.sqrt(S[0, 0]) * np.abs(U[:, 0]); H[0] = np.sqrt(S[0, 0]) * np.abs(V[0])
    for i in range(1, c["k"]):
        u, v, s = U[:, i], V[i], S[i, i]
        up, un, vp, vn = np.maximum(u, 0), np.maximum(-u, 0), np.maximum(v, 0), np.maximum(-v, 0)
        a, b = (up, vp) if np.linalg.norm(up) * np.linalg.norm(vp) >= np.linalg.norm(un) * np.linalg.norm(vn) else (un, vn)
        s2 = s * np.linalg.norm(a) * np.linalg.norm(b)
        W[:, i] = np.sqrt(s2) * a / np.linalg.norm(a)
        H[i] = np.sqrt(s2) * b / np.linalg.norm(b)
    assert rel(W, g["W"]) < 1e-9 and rel(H, g["H"]) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(cases.NNDSVD_CASES))
def test_gpu_nndsvd_matches_reference(name, golden_dir):
    c = cases.NNDSVD_CASES[name]
    g = np.load(os.path.join(golden_dir, "%s.npz" % name))
    X = cases.build_nndsvd(name)
    m = pymf_b200.NNDSVD(X, num_bases=c["k"])
    m.factorize()
    assert m.W.shape == g["W"].shape and m.H.shape == g["H"].shape and m.ferr.shape == (1,)
    assert np.max(np.abs(m.singular_values - g["sigma"][:c["k"]]) / g["sigma"][:c["k"]]) < 1e-5
    rw, rh = rel(m.W, g["W"]), rel(m.H, g["H"])
    print(name, "sweeps", m.svd_sweeps, "rel W %.2e rel H %.2e" % (rw, rh))
    assert rw < TOL_WH and rh < TOL_WH, (rw, rh)
    assert abs(m.ferr[0] - g["ferr"][0]) / g["ferr"][0] < TOL_FERR
    assert m.W.min() >= 0 and m.H.min() >= 0
    # ... and it feeds the multiplicative updates (class docstring of the reference, pymf/nndsvd.py:56-66)
    f = pymf_b200.NMF(X, num_bases=c["k"])
    f.W, f.H = m.W, m.H
    f.factorize(niter=10)
    assert rel(f.W, g["W_nmf10"]) < TOL_WH and rel(f.H, g["H_nmf10"]) < TOL_WH
    assert np.max(np.abs(f.ferr - g["ferr_nmf10"]) / g["ferr_nmf10"]) < TOL_FERR
