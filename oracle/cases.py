"""Seeded parity cases shared by make_golden.py, the tests, smoke() and bench.py.

TEST INFRASTRUCTURE ONLY.  ``build(name)`` regenerates (X, W0, H0) from the
recipe; the matching reference outputs live in tests/golden/traj_<name>.npz.
"""
import numpy as np

from .nmf_oracle import gen_matrix

REF_TEST_SEED = 400401          # tests/test_pymf.py:32
REF_TEST_INIT_SEED = 1234       # seed set right before the lazy W/H init in our fixtures


def ref_test_matrix():
    """tests/test_pymf.py:32-33."""
    np.random.seed(REF_TEST_SEED)
    return np.random.random((3, 50)) + 2.0


# kind "np":   X = np.random.random (float64) after np.random.seed(seed); W0 then H0 follow
#              from the same stream (SURVEY 8d, cfg1 recipe).
# kind "hash": X/W0/H0 from the counter hash (float32 values), seeds seed, seed+1, seed+2 -
#              the generator the device implements for the big synthetic configs.
CASES = {
    # BASELINE.json configs[0]
    "cfg1": dict(kind="np", seed=0, d=1000, n=500, k=10, niter=100, keep=[1, 10, 50, 100]),
    # ragged / odd shapes, nothing a multiple of anything
    "ragged": dict(kind="np", seed=7, d=37, n=201, k=5, niter=30, keep=[1, 2, 30]),
    "tiny": dict(kind="np", seed=11, d=3, n=7, k=2, niter=10, keep=[1, 10]),
    "k1": dict(kind="np", seed=13, d=50, n=64, k=1, niter=5, keep=[5]),
    # column prefix of cfg2 (d=4096, k=32) with the device generator
    "cfg2_prefix": dict(kind="hash", seed=1234, d=4096, n=2048, k=32, niter=5, keep=[1, 5],
                        store32=True),
    # k = 128 (cfg3's k), tensor-path shape, small n/d
    "k128": dict(kind="hash", seed=77, d=512, n=640, k=128, niter=5, keep=[1, 5], store32=True),
    # k not a multiple of 16, d not a multiple of 128
    "k40": dict(kind="hash", seed=99, d=300, n=1000, k=40, niter=8, keep=[1, 8], store32=True),
    # the same cfg2 prefix followed for cfg2's full 200 iterations (north_star: <= 1e-4 per iteration)
    "cfg2_prefix_200": dict(kind="hash", seed=1234, d=4096, n=2048, k=32, niter=200, keep=[1, 50, 100, 200],
                            store32=True),
    # cfg3's d and k (16384 rows = 512 MMA stages = 16 TMEM segments per column tile), cfg5's d and k.
    # W snapshots keep every 8th row (w_stride) so that the fixtures stay ~1 MB; normW / normH are of the
    # full matrices, and the GPU tests also run the oracle live on the full factors.
    "cfg3_d": dict(kind="hash", seed=301, d=16384, n=2048, k=128, niter=4, keep=[1, 4], store32=True, w_stride=8),
    "cfg5_d": dict(kind="hash", seed=501, d=32768, n=1024, k=64, niter=4, keep=[1, 4], store32=True, w_stride=8),
    # cross-rank parity case of bench.py (N > 1): 2048 columns = 256 per rank at N = 8, tensor-path shape
    "scale_k128": dict(kind="hash", seed=801, d=1024, n=2048, k=128, niter=6, keep=[6], store32=True),
}


# BNMF parity cases (pymf/bnmf.py).  kind "bin": X is a planted BINARY matrix (W* H* > 0 with sparse
# binary factors, the data BNMF is made for); W0/H0 ~ U[0,1) from the same numpy stream.
BNMF_CASES = {
    "bnmf_bin": dict(kind="bin", seed=21, d=64, n=200, k=6, niter=30, keep=[1, 10, 30]),
    "bnmf_ragged": dict(kind="np", seed=23, d=37, n=201, k=5, niter=20, keep=[1, 20]),
    "bnmf_tc": dict(kind="hash", seed=55, d=512, n=640, k=32, niter=12, keep=[1, 12], store32=True),
    "bnmf_k40": dict(kind="hash", seed=57, d=300, n=1000, k=40, niter=10, keep=[1, 10], store32=True),
}


# SNMF parity cases (pymf/snmf.py).  kind "npc" / "hashc": the data is CENTERED (X - 0.5, signed entries), which
# is what semi-NMF is for; W0/H0 ~ U[0,1).
SNMF_CASES = {
    "snmf_signed": dict(kind="npc", seed=31, d=40, n=150, k=4, niter=15, keep=[1, 5, 15]),
    "snmf_ragged": dict(kind="np", seed=33, d=37, n=201, k=5, niter=12, keep=[1, 12]),
    "snmf_tc": dict(kind="hashc", seed=61, d=512, n=640, k=32, niter=8, keep=[1, 8], store32=True),
    "snmf_k40": dict(kind="hash", seed=63, d=300, n=1000, k=40, niter=6, keep=[1, 6], store32=True),
}


# NNDSVD parity cases (pymf/nndsvd.py).  kind "lowrank": X = W* H* + 0.02 U with W*, H* >= 0 whose columns / rows are
# scaled by a geometric sequence, so that the leading singular values are well separated (singular VECTORS are only
# defined - and only comparable between an exact SVD and an iterative one - where the gaps are not tiny).
NNDSVD_CASES = {
    "nndsvd_small": dict(kind="lowrank", seed=41, d=60, n=90, k=4, rank=6),
    "nndsvd_wide": dict(kind="lowrank", seed=43, d=300, n=2000, k=8, rank=12),
    "nndsvd_tall": dict(kind="lowrank", seed=45, d=700, n=400, k=6, rank=10),       # rows > cols: the reference's _left_svd
    "nndsvd_tc": dict(kind="lowrank", seed=47, d=1024, n=4096, k=16, rank=24),      # tensor-path shape of the NMF that follows
}


def build_nndsvd(name):
    c = NNDSVD_CASES[name]
    rng = np.random.RandomState(c["seed"])
    d, n, r = c["d"], c["n"], c["rank"]
    scale = 0.75 ** np.arange(r)
    Ws = rng.random_sample((d, r)) * (rng.random_sample((d, r)) < 0.5) * scale[None, :]
    Hs = rng.random_sample((r, n)) * (rng.random_sample((r, n)) < 0.5)
    return Ws.dot(Hs) + 0.02 * rng.random_sample((d, n))


def build(name):
    c = CASES[name] if name in CASES else (BNMF_CASES[name] if name in BNMF_CASES else SNMF_CASES[name])
    d, n, k = c["d"], c["n"], c["k"]
    if c["kind"] == "bin":
        np.random.seed(c["seed"])
        Ws = (np.random.random((d, k)) < 0.3).astype(np.float64)
        Hs = (np.random.random((k, n)) < 0.3).astype(np.float64)
        X = (Ws.dot(Hs) > 0).astype(np.float64)
        W0 = np.random.random((d, k))
        H0 = np.random.random((k, n))
    elif c["kind"] in ("np", "npc"):
        np.random.seed(c["seed"])
        X = np.random.random((d, n)) - (0.5 if c["kind"] == "npc" else 0.0)
        W0 = np.random.random((d, k))
        H0 = np.random.random((k, n))
    else:
        X = gen_matrix(c["seed"], d, n)
        if c["kind"] == "hashc":
            X = X - np.float32(0.5)
        W0 = gen_matrix(c["seed"] + 1, d, k).astype(np.float64)
        H0 = gen_matrix(c["seed"] + 2, k, n).astype(np.float64)
    return X, W0, H0
