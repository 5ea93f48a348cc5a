"""Column-sharded N > 1 path on CPU: world_size 2 over gloo, host logic of pymf_b200.NMF
(shard bookkeeping, global n in converged(), replicated W, sharded H, identical RNG stream)
with the oracle-backed engine double standing in for the GPU engine."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from oracle import nmf_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, d, n, k, niter, out, cls="NMF"):
    import torch.distributed as dist
    import pymf_b200
    from tests._fake_engine import FakeEngine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pymf_b200.NMF._engine_factory = FakeEngine
        X = np.random.RandomState(42).random_sample((d, n))
        bounds = [int(round(n * r / float(world))) for r in range(world + 1)]
        lo, hi = bounds[rank], bounds[rank + 1]
        np.random.seed(9)                                   # same stream on every rank
        m = getattr(pymf_b200, cls)(X[:, lo:hi], num_bases=k, process_group=True)
        assert m._num_samples == n and m._col0 == lo
        m.factorize(niter=niter)
        np.savez(os.path.join(out, "r%d.npz" % rank), W=m.W, H=m.H, ferr=m.ferr, lo=lo, hi=hi)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [64, 201])
def test_two_rank_column_sharding_matches_single_process(tmp_path, n):
    d, k, niter, world = 17, 4, 12, 2
    mp.spawn(_worker, args=(world, _free_port(), d, n, k, niter, str(tmp_path)), nprocs=world, join=True)
    X = np.random.RandomState(42).random_sample((d, n))
    np.random.seed(9)
    W, H = O.init_wh(d, n, k)
    ferr = O.factorize(X, W, H, niter=niter)
    parts = [np.load(os.path.join(str(tmp_path), "r%d.npz" % r)) for r in range(world)]
    for p in parts:
        np.testing.assert_allclose(p["W"], W, rtol=1e-9)               # replicated
        np.testing.assert_allclose(p["H"], H[:, int(p["lo"]):int(p["hi"])], rtol=1e-9)
        np.testing.assert_allclose(p["ferr"], ferr, rtol=1e-9)          # global error
    np.testing.assert_array_equal(parts[0]["W"], parts[1]["W"])         # bit-identical replicas


@pytest.mark.parametrize("cls", ["BNMF", "SNMF"])
def test_two_rank_column_sharding_of_the_variants(tmp_path, cls):
    """BNMF / SNMF shard exactly like NMF: their W updates consume the same summed X H^T / H H^T partials."""
    d, n, k, niter, world = 17, 101, 4, 9, 2
    mp.spawn(_worker, args=(world, _free_port(), d, n, k, niter, str(tmp_path), cls), nprocs=world, join=True)
    X = np.random.RandomState(42).random_sample((d, n))
    np.random.seed(9)
    W, H = O.init_wh(d, n, k)
    if cls == "BNMF":
        ferr = O.bnmf_factorize(X, W, H, niter=niter)
    else:
        W, ferr = O.snmf_factorize(X, W, H, niter=niter)
    parts = [np.load(os.path.join(str(tmp_path), "r%d.npz" % r)) for r in range(world)]
    for p in parts:
        np.testing.assert_allclose(p["W"], W, rtol=1e-8)
        np.testing.assert_allclose(p["H"], H[:, int(p["lo"]):int(p["hi"])], rtol=1e-8, atol=1e-300)
        np.testing.assert_allclose(p["ferr"], ferr, rtol=1e-9)
    np.testing.assert_array_equal(parts[0]["W"], parts[1]["W"])


def _worker_unseeded(rank, world, port, d, n, k, niter, out):
    import torch.distributed as dist
    import pymf_b200
    from tests._fake_engine import FakeEngine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pymf_b200.NMF._engine_factory = FakeEngine
        X = np.random.RandomState(42).random_sample((d, n))
        lo, hi = (0, n // 2) if rank == 0 else (n // 2, n)
        np.random.seed(1000 + rank)                         # every rank on its OWN random stream
        m = pymf_b200.NMF(X[:, lo:hi], num_bases=k, process_group=True)
        m.factorize(niter=niter)                            # lazy init: each rank draws a different W
        w_lazy = m.W.copy()
        m.W = np.random.random((d, k))                      # host-assigned, different on every rank
        m.factorize(niter=niter)
        np.savez(os.path.join(out, "u%d.npz" % rank), W1=w_lazy, W=m.W, H=m.H, ferr=m.ferr, lo=lo, hi=hi)
    finally:
        dist.destroy_process_group()


def test_unseeded_ranks_still_hold_one_replicated_w(tmp_path):
    """Without a common numpy seed every rank draws / assigns a different W; the host layer broadcasts rank 0's
    before the first upload, so the replicas are bit-identical and the reported error is the true ||X - W H||."""
    d, n, k, niter, world = 17, 90, 4, 15, 2
    mp.spawn(_worker_unseeded, args=(world, _free_port(), d, n, k, niter, str(tmp_path)), nprocs=world, join=True)
    X = np.random.RandomState(42).random_sample((d, n))
    p0, p1 = [np.load(os.path.join(str(tmp_path), "u%d.npz" % r)) for r in range(world)]
    np.testing.assert_array_equal(p0["W1"], p1["W1"])
    np.testing.assert_array_equal(p0["W"], p1["W"])
    np.testing.assert_array_equal(p0["ferr"], p1["ferr"])
    H = np.concatenate([p0["H"], p1["H"]], axis=1)
    true = O.frobenius_norm(X, p0["W"], H)
    assert abs(p0["ferr"][-1] - true) / true < 1e-9
