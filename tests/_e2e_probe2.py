"""GPU experiment (not a test): stage timing of the bench's e2e leg (NMF with pinned host buffers),
1-D vs 2-D DMA, and the effect of a previous big engine in the same process."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pymf_b200  # noqa: E402
from pymf_b200.engine import Engine  # noqa: E402


def timed(cls, name, log):
    orig = getattr(cls, name)

    def wrap(self, *a, **kw):
        t0 = time.perf_counter()
        r = orig(self, *a, **kw)
        log.append((name, time.perf_counter() - t0))
        return r
    setattr(cls, name, wrap)


def main():
    d, n, k = 4096, 262144, 32
    log = []
    for nm in ("__init__", "upload_x", "set_w", "set_h", "run", "get_w", "get_h"):
        timed(Engine, nm, log)
    if "--with-torch" in sys.argv:
        import torch
        torch.cuda.set_device(0)
        torch.cuda.synchronize()
    if "--prior-engine" in sys.argv:
        e = Engine(d, n, k)
        e.gen_x(1); e.gen_w(2); e.gen_h(3); e.enqueue(5); e.sync(); e.close()
        del log[:]
    rng = np.random.default_rng(0)
    Xp = pymf_b200.pinned_empty((d, n), np.float32); rng.random(out=Xp, dtype=np.float32)
    Xq = pymf_b200.pinned_empty((d, n + 32), np.float32); Xq[:, :n] = Xp
    W0 = pymf_b200.pinned_empty((d, k), np.float64); rng.random(out=W0)
    H0 = pymf_b200.pinned_empty((k, n), np.float64); rng.random(out=H0)
    for rep, X in enumerate((Xp, Xq[:, :n], Xp, Xq[:, :n])):
        del log[:]
        t0 = time.perf_counter()
        m = pymf_b200.NMF(X, num_bases=k)
        m.W, m.H = W0, H0
        m._sync_to_device()
        t1 = time.perf_counter()
        m.factorize(niter=20)
        t2 = time.perf_counter()
        _ = (m.W, m.H, m.ferr)
        t3 = time.perf_counter()
        print("rep %d (%s): total %.4f  sync_to_device %.4f  factorize %.4f  download %.4f | %s" % (
            rep, "1-D" if X is Xp else "2-D", t3 - t0, t1 - t0, t2 - t1, t3 - t2,
            "  ".join("%s %.4f" % (a, b) for a, b in log)), flush=True)
        del m


if __name__ == "__main__":
    main()
