"""Generate tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference/pymf/nmf.py, loaded by path - see ref_loader.py).

TEST INFRASTRUCTURE ONLY.  Run in the build container (the GPU box has no
reference checkout):   python -m oracle.make_golden

Each fixture stores the seeds/recipe that regenerate the inputs plus the
reference's outputs (W, H, ferr), so the fixtures stay small.  Inputs are
rebuilt by ``oracle.cases`` from the recipe at test time.
"""
import os
import sys

import numpy as np

from . import cases
from .ref_loader import load_reference_bnmf, load_reference_nmf, load_reference_nndsvd, load_reference_snmf

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def run_steps(ref, X, W0, H0, k, niter, keep, cls="NMF"):
    """Single-step the reference (factorize(niter=1) never triggers the early stop,
    SURVEY 3.4) and record ferr per iteration and W/H at the iterations in keep."""
    m = getattr(ref, cls)(X, num_bases=k)
    m.W = W0.copy()
    m.H = H0.copy()
    ferr = np.zeros(niter)
    normW = np.zeros(niter)
    normH = np.zeros(niter)
    snaps = {}
    for i in range(niter):
        m.factorize(niter=1)
        ferr[i] = m.ferr[0]
        normW[i] = np.linalg.norm(m.W)
        normH[i] = np.linalg.norm(m.H)
        if (i + 1) in keep:
            snaps["W_%d" % (i + 1)] = m.W.copy()
            snaps["H_%d" % (i + 1)] = m.H.copy()
    return ferr, normW, normH, snaps


def main():
    ref = load_reference_nmf()
    if ref is None:
        sys.exit("no reference checkout found (set PYMF_REF)")
    os.makedirs(OUT, exist_ok=True)

    # ---- 1. the reference's own test case (tests/test_pymf.py:32-33,69,84-95) ----
    A = cases.ref_test_matrix()
    np.random.seed(cases.REF_TEST_INIT_SEED)
    m = ref.NMF(A, num_bases=4)
    m.factorize(niter=20)                       # lazy init: W then H from np.random
    out = {"W_20": m.W.copy(), "H_20": m.H.copy(), "ferr_20": m.ferr.copy()}
    m.factorize(compute_h=False)                # :92
    out.update(W_a=m.W.copy(), H_a=m.H.copy(), ferr_a=m.ferr.copy())
    m.factorize(compute_w=False)                # :93
    out.update(W_b=m.W.copy(), H_b=m.H.copy(), ferr_b=m.ferr.copy())
    m.factorize(compute_err=False)              # :94 (ferr untouched)
    out.update(W_c=m.W.copy(), H_c=m.H.copy(), ferr_c=m.ferr.copy())
    m.factorize(niter=20)                       # :95 warm start
    out.update(W_d=m.W.copy(), H_d=m.H.copy(), ferr_d=m.ferr.copy())
    np.savez(os.path.join(OUT, "ref_test_3x50.npz"), **out)
    assert out["ferr_20"][-1] / 53 < 0.1

    # ---- 2. early stop / truncation on the same matrix ----
    np.random.seed(cases.REF_TEST_INIT_SEED)
    m = ref.NMF(A, num_bases=4)
    m.factorize(niter=100000)
    np.savez(os.path.join(OUT, "ref_conv_3x50.npz"), W=m.W.copy(), H=m.H.copy(),
             ferr_len=np.int64(len(m.ferr)), ferr_tail=m.ferr[-8:].copy(),
             ferr_head=m.ferr[:8].copy())
    print("early stop: len(ferr) =", len(m.ferr))

    # ---- 3. trajectories on the parity cases ----
    only = set(sys.argv[2:]) if len(sys.argv) > 2 and sys.argv[1] == "only" else None
    for name, c in cases.CASES.items():
        if not c.get("golden", True) or (only is not None and name not in only):
            continue
        X, W0, H0 = cases.build(name)
        ferr, nW, nH, snaps = run_steps(ref, X.astype(np.float64), W0, H0, c["k"], c["niter"],
                                        set(c["keep"]))
        if c.get("w_stride", 1) > 1:
            snaps = {k_: (v[::c["w_stride"]] if k_.startswith("W_") else v) for k_, v in snaps.items()}
        if c.get("store32", False):
            snaps = {k_: v.astype(np.float32) for k_, v in snaps.items()}
        np.savez(os.path.join(OUT, "traj_%s.npz" % name), ferr=ferr, normW=nW, normH=nH, **snaps)
        print(name, "ferr", ferr[0], "->", ferr[-1])


def main_bnmf():
    """BNMF fixtures: the reference's hooks stepped exactly as NMF.factorize steps them
    (pymf/nmf.py:182-190) after BNMF.factorize's lamb initialisation (pymf/bnmf.py:117-118), plus one
    whole factorize() call per case for the end state / ferr vector / final lamb."""
    refb = load_reference_bnmf()
    if refb is None:
        sys.exit("no reference checkout found (set PYMF_REF)")
    # the reference's own BNMF test vector (tests/test_pymf.py:80,84-95): np.round(A - 2.0), k = 4, niter = 20,
    # then the flag combinations and a warm restart
    A = np.round(cases.ref_test_matrix() - 2.0)
    np.random.seed(cases.REF_TEST_INIT_SEED)
    m = refb.BNMF(A, num_bases=4)
    m.factorize(niter=20)
    out = {"W_20": m.W.copy(), "H_20": m.H.copy(), "ferr_20": m.ferr.copy()}
    assert out["ferr_20"][-1] / 53 < 0.1                     # the reference's bound, :86-88
    m.factorize(compute_h=False, niter=20)                   # :92
    out.update(W_a=m.W.copy(), H_a=m.H.copy(), ferr_a=m.ferr.copy())
    m.factorize(compute_w=False, niter=20)                   # :93
    out.update(W_b=m.W.copy(), H_b=m.H.copy(), ferr_b=m.ferr.copy())
    m.factorize(compute_err=False, niter=20)                 # :94
    out.update(W_c=m.W.copy(), H_c=m.H.copy(), ferr_c=m.ferr.copy())
    m.factorize(niter=20)                                    # :95
    out.update(W_d=m.W.copy(), H_d=m.H.copy(), ferr_d=m.ferr.copy(), lam_d=np.array([m._lamb_W, m._lamb_H]))
    np.savez(os.path.join(OUT, "bnmf_ref_test_3x50.npz"), **out)
    print("bnmf ref test: ferr[-1]/53 =", out["ferr_20"][-1] / 53, "lens", [len(out[k_]) for k_ in ("ferr_20", "ferr_a", "ferr_b", "ferr_c", "ferr_d")])

    for name, c in cases.BNMF_CASES.items():
        X, W0, H0 = cases.build(name)
        X = X.astype(np.float64)
        niter, keep = c["niter"], set(c["keep"])
        m = refb.BNMF(X, num_bases=c["k"])
        m.W, m.H = W0.copy(), H0.copy()
        m._lamb_W = m._lamb_H = 1.0 / niter
        ferr = np.zeros(niter)
        snaps = {}
        for i in range(niter):
            m.update_w()
            m.update_h()
            ferr[i] = m.frobenius_norm()
            if (i + 1) in keep:
                snaps["W_%d" % (i + 1)] = m.W.copy()
                snaps["H_%d" % (i + 1)] = m.H.copy()
        lam_steps = np.array([m._lamb_W, m._lamb_H])
        f = refb.BNMF(X, num_bases=c["k"])
        f.W, f.H = W0.copy(), H0.copy()
        f.factorize(niter=niter)
        whole = {"Wf": f.W.copy(), "Hf": f.H.copy()}
        if c.get("store32", False):
            snaps = {k_: v.astype(np.float32) for k_, v in snaps.items()}
            whole = {k_: v.astype(np.float32) for k_, v in whole.items()}
        np.savez(os.path.join(OUT, "%s.npz" % name), ferr=ferr, lam_steps=lam_steps, ferr_whole=f.ferr.copy(),
                 lam_whole=np.array([f._lamb_W, f._lamb_H]), **snaps, **whole)
        print(name, "ferr", ferr[0], "->", ferr[-1], "len(ferr_whole)", len(f.ferr), "lam", f._lamb_H)


def main_snmf():
    """SNMF fixtures from the unmodified pymf/snmf.py: the reference's own test vector
    (tests/test_pymf.py:77,84-95) and stepped trajectories of the seeded cases."""
    refs = load_reference_snmf()
    if refs is None:
        sys.exit("no reference checkout found (set PYMF_REF)")
    A = cases.ref_test_matrix()
    np.random.seed(cases.REF_TEST_INIT_SEED)
    m = refs.SNMF(A, num_bases=4)
    m.factorize(niter=20)
    out = {"W_20": m.W.copy(), "H_20": m.H.copy(), "ferr_20": m.ferr.copy()}
    assert out["ferr_20"][-1] / 53 < 0.1
    m.factorize(compute_h=False, niter=20)
    out.update(W_a=m.W.copy(), H_a=m.H.copy(), ferr_a=m.ferr.copy())
    m.factorize(compute_w=False, niter=20)
    out.update(W_b=m.W.copy(), H_b=m.H.copy(), ferr_b=m.ferr.copy())
    m.factorize(compute_err=False, niter=20)
    out.update(W_c=m.W.copy(), H_c=m.H.copy(), ferr_c=m.ferr.copy())
    m.factorize(niter=20)
    out.update(W_d=m.W.copy(), H_d=m.H.copy(), ferr_d=m.ferr.copy())
    np.savez(os.path.join(OUT, "snmf_ref_test_3x50.npz"), **out)
    print("snmf ref test: ferr[-1]/53 =", out["ferr_20"][-1] / 53, "lens",
          [len(out[k_]) for k_ in ("ferr_20", "ferr_a", "ferr_b", "ferr_c", "ferr_d")])
    for name, c in cases.SNMF_CASES.items():
        X, W0, H0 = cases.build(name)
        ferr, nW, nH, snaps = run_steps(refs, X.astype(np.float64), W0, H0, c["k"], c["niter"], set(c["keep"]),
                                        cls="SNMF")
        if c.get("store32", False):
            snaps = {k_: v.astype(np.float32) for k_, v in snaps.items()}
        np.savez(os.path.join(OUT, "%s.npz" % name), ferr=ferr, normW=nW, normH=nH, **snaps)
        print(name, "ferr", ferr[0], "->", ferr[-1])


def main_nndsvd():
    """NNDSVD fixtures from the unmodified pymf/nndsvd.py + pymf/svd.py: W, H, ferr of NNDSVD(X, k).factorize(), the
    leading singular values, and the first 10 iterations of the reference NMF warm-started from them."""
    refn = load_reference_nndsvd()
    ref = load_reference_nmf()
    if refn is None:
        sys.exit("no reference checkout found (set PYMF_REF)")
    import sys as _sys
    svd_mod = _sys.modules["_pymf_ref_pkg.svd"]
    for name, c in cases.NNDSVD_CASES.items():
        X = cases.build_nndsvd(name)
        m = refn.NNDSVD(X, num_bases=c["k"])
        m.factorize()
        sv = svd_mod.SVD(X)
        sv.factorize()
        f = ref.NMF(X, num_bases=c["k"])
        f.W, f.H = m.W.copy(), m.H.copy()
        f.factorize(niter=10)
        np.savez(os.path.join(OUT, "%s.npz" % name), W=m.W, H=m.H, ferr=m.ferr, sigma=np.diag(sv.S)[:c["k"] + 4].copy(),
                 W_nmf10=f.W, H_nmf10=f.H, ferr_nmf10=f.ferr)
        print(name, "ferr", m.ferr, "sigma", np.diag(sv.S)[:c["k"] + 1], "nmf ferr", f.ferr[0], "->", f.ferr[-1])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "nndsvd":
        main_nndsvd()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "bnmf":
        main_bnmf()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "snmf":
        main_snmf()
        sys.exit(0)
    main()
