"""CPU oracle: numpy restatement of pymf's NMF multiplicative-update path.

TEST INFRASTRUCTURE ONLY.  The product (``pymf_b200``) never imports this file;
it is the checker for the CUDA path and the "port" CPU baseline of ``bench.py``.

Pinning: the reference's own tests hold no golden vectors for this path
(``tests/test_pymf.py:69,86-88`` is a loose smoke bound).  The oracle is therefore
pinned against outputs of the reference itself: ``oracle/make_golden.py`` runs the
unmodified ``/root/reference/pymf/nmf.py`` (loaded by path, see ``ref_loader.py``)
and commits its trajectories under ``tests/golden/``; ``tests/test_oracle.py``
requires this restatement to reproduce them to 1e-12 relative.

Every function cites the reference lines it follows (paths relative to the
reference checkout).  float64 throughout, same operation order as the reference.
"""
import numpy as np

EPS_DENOM = 10 ** -9          # pymf/nmf.py:124,130 (added to the denominator only)
EPS_CONV = 10 ** -8           # pymf/nmf.py:69  (NMF._EPS, convergence threshold)
SENTINEL = -123456            # pymf/nmf.py:112


def update_h(X, W, H):
    """In-place H update.  pymf/nmf.py:122-126."""
    H2 = np.dot(np.dot(W.T, W), H) + EPS_DENOM      # :124
    H *= np.dot(W.T, X)                             # :125
    H /= H2                                         # :126
    return H


def update_w(X, W, H):
    """In-place W update (uses the H from before this iteration).  pymf/nmf.py:128-132."""
    W2 = np.dot(np.dot(W, H), H.T) + EPS_DENOM      # :130
    W *= np.dot(X, H.T)                             # :131
    W /= W2                                         # :132
    return W


def frobenius_norm(X, W, H):
    """||X - W H||_F.  pymf/nmf.py:100-114 (dense branch, :110)."""
    return np.sqrt(np.sum((X - np.dot(W, H)) ** 2))


def converged(ferr, i, num_samples):
    """pymf/nmf.py:134-139."""
    derr = np.abs(ferr[i] - ferr[i - 1]) / num_samples
    return bool(derr < EPS_CONV)


def init_wh(d, n, k):
    """Lazy initialisation order of factorize(): W first, then H, from numpy's
    global RNG.  pymf/nmf.py:116-120,173-177."""
    W = np.random.random((d, k))
    H = np.random.random((k, n))
    return W, H


def factorize(X, W, H, niter=1, compute_w=True, compute_h=True, compute_err=True,
              early_stop=True, record=None):
    """Driver loop.  pymf/nmf.py:141-202: per iteration W, then H, then error;
    convergence is only checked for i > 1 and drops entry i of ferr.

    W and H are float64 arrays updated in place.  Returns ferr (or None when
    compute_err is False).  ``record(i, W, H, ferr_i)`` is called after every
    iteration when given (parity harness).
    """
    ferr = np.zeros(niter) if compute_err else None        # :179-180
    for i in range(niter):                                 # :182
        if compute_w:
            update_w(X, W, H)                              # :183-184
        if compute_h:
            update_h(X, W, H)                              # :186-187
        if compute_err:
            ferr[i] = frobenius_norm(X, W, H)              # :189-190
        if record is not None:
            record(i, W, H, ferr[i] if compute_err else None)
        if early_stop and i > 1 and compute_err:           # :198-202
            if converged(ferr, i, X.shape[1]):
                ferr = ferr[:i]
                break
    return ferr


# --------------------------------------------------------------------------
# BNMF (pymf/bnmf.py): the same loop with a penalty that drives W and H towards {0, 1}.
# Pinned like the NMF functions: tests/golden/bnmf_*.npz come from the unmodified reference.
# --------------------------------------------------------------------------
LAMB_INCREASE_W = 1.1         # pymf/bnmf.py:75
LAMB_INCREASE_H = 1.1         # pymf/bnmf.py:76


def bnmf_update_h(X, W, H, lam):
    """In-place BNMF H update; lam = {"W": lamb_W, "H": lamb_H} grows here.  pymf/bnmf.py:78-85."""
    H1 = np.dot(W.T, X) + 3.0 * lam["H"] * (H ** 2)                                         # :79
    H2 = np.dot(np.dot(W.T, W), H) + 2 * lam["H"] * (H ** 3) + lam["H"] * H + EPS_DENOM     # :80
    H *= H1 / H2                                                                            # :81
    lam["W"] = LAMB_INCREASE_W * lam["W"]                                                   # :83
    lam["H"] = LAMB_INCREASE_H * lam["H"]                                                   # :84
    return H


def bnmf_update_w(X, W, H, lam):
    """In-place BNMF W update.  pymf/bnmf.py:86-89."""
    W1 = np.dot(X, H.T) + 3.0 * lam["W"] * (W ** 2)                                         # :87
    W2 = np.dot(W, np.dot(H, H.T)) + 2.0 * lam["W"] * (W ** 3) + lam["W"] * W + EPS_DENOM   # :88
    W *= W1 / W2                                                                            # :89
    return W


def bnmf_factorize(X, W, H, niter=10, compute_w=True, compute_h=True, compute_err=True,
                   early_stop=True, record=None, lam=None):
    """pymf/bnmf.py:91-123: lamb_W = lamb_H = 1/niter, then NMF.factorize's loop
    (pymf/nmf.py:182-202) over the overridden hooks.  Returns ferr (or None); the final
    penalty weights are left in ``lam`` when a dict is passed."""
    lam = {} if lam is None else lam
    lam["W"] = 1.0 / niter                                 # :117
    lam["H"] = 1.0 / niter                                 # :118
    ferr = np.zeros(niter) if compute_err else None
    for i in range(niter):
        if compute_w:
            bnmf_update_w(X, W, H, lam)
        if compute_h:
            bnmf_update_h(X, W, H, lam)
        if compute_err:
            ferr[i] = frobenius_norm(X, W, H)
        if record is not None:
            record(i, W, H, ferr[i] if compute_err else None)
        if early_stop and i > 1 and compute_err:
            if converged(ferr, i, X.shape[1]):
                ferr = ferr[:i]
                break
    return ferr


# --------------------------------------------------------------------------
# SNMF (pymf/snmf.py): semi-NMF - X and W signed, H >= 0.  Pinned by tests/golden/snmf_*.npz.
# --------------------------------------------------------------------------
def snmf_update_w(X, W, H):
    """Returns the NEW W (the reference rebinds self.W).  pymf/snmf.py:67-70."""
    W1 = np.dot(X, H.T)                                     # :68
    W2 = np.dot(H, H.T)                                     # :69
    return np.dot(W1, np.linalg.inv(W2))                    # :70


def snmf_update_h(X, W, H):
    """In-place H update.  pymf/snmf.py:72-90."""
    def separate_positive(m):                               # :73-74
        return (np.abs(m) + m) / 2.0

    def separate_negative(m):                               # :76-77
        return (np.abs(m) - m) / 2.0

    XW = np.dot(X.T, W)                                     # :79
    WW = np.dot(W.T, W)                                     # :81
    WW_pos = separate_positive(WW)                          # :82
    WW_neg = separate_negative(WW)                          # :83
    XW_pos = separate_positive(XW)                          # :85
    H1 = (XW_pos + np.dot(H.T, WW_neg)).T                   # :86
    XW_neg = separate_negative(XW)                          # :88
    H2 = (XW_neg + np.dot(H.T, WW_pos)).T + EPS_DENOM       # :89
    H *= np.sqrt(H1 / H2)                                   # :90
    return H


def snmf_factorize(X, W, H, niter=1, compute_w=True, compute_h=True, compute_err=True,
                   early_stop=True, record=None):
    """NMF.factorize's loop (pymf/nmf.py:182-202) over SNMF's hooks.  Returns (W, ferr): W is
    rebound by update_w, H is updated in place."""
    ferr = np.zeros(niter) if compute_err else None
    for i in range(niter):
        if compute_w:
            W = snmf_update_w(X, W, H)
        if compute_h:
            snmf_update_h(X, W, H)
        if compute_err:
            ferr[i] = frobenius_norm(X, W, H)
        if record is not None:
            record(i, W, H, ferr[i] if compute_err else None)
        if early_stop and i > 1 and compute_err:
            if converged(ferr, i, X.shape[1]):
                ferr = ferr[:i]
                break
    return W, ferr


# --------------------------------------------------------------------------
# NNDSVD (pymf/nndsvd.py) and the dense SVD it calls (pymf/svd.py).  Pinned by tests/golden/nndsvd_*.npz, which
# come from the unmodified reference files.
# --------------------------------------------------------------------------
SVD_EPS = 10 ** -8            # pymf/svd.py:74


def svd_dense(X):
    """U, S (diagonal matrix), V with X = U S V.  pymf/svd.py:111-158,237-246: eigh of X X^T when rows <= cols
    (_right_svd), of X^T X otherwise (_left_svd); eigenvalues <= 1e-8 are dropped, the rest sorted descending."""
    rows, cols = X.shape
    if rows > cols:                                              # _left_svd, :137-158
        values, v_vectors = np.linalg.eigh(np.dot(X.T, X))       # :138-139
        v_vectors = v_vectors[:, values > SVD_EPS]               # :142
        values = values[values > SVD_EPS]                        # :143
        idx = np.argsort(values)[::-1]                           # :147
        values = values[idx]
        S = np.diag(np.sqrt(values))                             # :151
        S_inv = np.diag(1.0 / np.sqrt(values))                   # :154
        Vtmp = v_vectors[:, idx]                                 # :156
        U = np.dot(np.dot(X, Vtmp), S_inv)                       # :158
        V = Vtmp.T
    else:                                                        # _right_svd, :112-134
        values, u_vectors = np.linalg.eigh(np.dot(X, X.T))       # :113-114
        u_vectors = u_vectors[:, values > SVD_EPS]               # :117
        values = values[values > SVD_EPS]                        # :118
        idx = np.argsort(values)                                 # :121
        values = values[idx[::-1]]                               # :122
        U = u_vectors[:, idx[::-1]]                              # :125
        S = np.diag(np.sqrt(values))                             # :128
        S_inv = np.diag(np.sqrt(values) ** -1)                   # :131
        V = np.dot(S_inv, np.dot(U.T, X))                        # :134
    return U, S, V


def nndsvd(X, k):
    """W (d x k), H (k x n) of NNDSVD.update_w.  pymf/nndsvd.py:79-108 (init_w / init_h zeros, :70-74)."""
    d, n = X.shape
    W = np.zeros((d, k))
    H = np.zeros((k, n))
    U, S, V = svd_dense(X)                                       # :80-83
    W[:, 0] = np.sqrt(S[0, 0]) * np.abs(U[:, 0])                 # :87
    H[0, :] = np.sqrt(S[0, 0]) * np.abs(V[0, :].T)               # :90
    for i in range(1, k):                                        # :92
        Tmp = np.dot(U[:, i:i + 1] * S[i, i], V[i:i + 1, :])     # :94
        Tmp = np.where(Tmp < 0, 0.0, Tmp)                        # :97
        u, s, v = svd_dense(Tmp)                                 # :100-102
        W[:, i] = np.sqrt(s[0, 0]) * np.abs(u[:, 0])             # :105
        H[i, :] = np.sqrt(s[0, 0]) * np.abs(v[0, :].T)           # :108
    return W, H


# --------------------------------------------------------------------------
# Synthetic inputs shared by tests, smoke() and bench.py.
# The device generator in pymf_b200/csrc/pymfb.cu (k_gen_uniform) implements the
# same integer hash so that any shard / tile regenerates bit-identically.
# --------------------------------------------------------------------------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def hash_uniform(seed, idx):
    """U[0,1) float32 from a splitmix64-style hash of (seed, element index).

    idx: uint64 array of global element indices.  Returns float32 with 24 random
    mantissa bits: (h >> 40) * 2**-24.
    """
    with np.errstate(over="ignore"):
        z = idx.astype(np.uint64) + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return ((z >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24))


def gen_matrix(seed, rows, cols, ld=None, col0=0, ncols=None):
    """rows x ncols block (columns col0..col0+ncols) of the synthetic rows x cols
    matrix whose element (r, c) is hash_uniform(seed, r*ld + c); ld defaults to cols."""
    ld = cols if ld is None else ld
    ncols = cols - col0 if ncols is None else ncols
    r = np.arange(rows, dtype=np.uint64)[:, None]
    c = np.arange(col0, col0 + ncols, dtype=np.uint64)[None, :]
    return hash_uniform(seed, r * np.uint64(ld) + c)
