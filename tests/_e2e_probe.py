"""GPU experiment (not a test): where does the end-to-end time of NMF(X_host).factorize() go?"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pymf_b200  # noqa: E402


def main():
    d, n, k = 4096, 262144, 32
    rng = np.random.default_rng(0)
    X = rng.random((d, n), dtype=np.float32)
    W0 = rng.random((d, k))
    H0 = rng.random((k, n))
    Xp = pymf_b200.pinned_empty((d, n), np.float32); Xp[...] = X
    W0p = pymf_b200.pinned_empty((d, k), np.float64); W0p[...] = W0
    H0p = pymf_b200.pinned_empty((k, n), np.float64); H0p[...] = H0
    for rep in range(3):
        t = [time.perf_counter()]
        eng = pymf_b200.Engine(d, n, k, device=0); t.append(time.perf_counter())
        eng.upload_x(Xp); t.append(time.perf_counter())
        eng.set_w(W0p); eng.set_h(H0p); t.append(time.perf_counter())
        eng.run(20); t.append(time.perf_counter())
        eng.get_w(out=W0p); eng.get_h(out=H0p); t.append(time.perf_counter())
        eng.close(); t.append(time.perf_counter())
        names = ["create", "upload_x", "set_w/h", "run20", "get_w/h", "close"]
        print("PINNED rep %d: " % rep + "  ".join("%s %.4f" % (nm, t[i + 1] - t[i]) for i, nm in enumerate(names)),
              " upload GB/s %.1f direct=%s" % (X.nbytes / 1e9 / (t[2] - t[1]), eng.last_upload_pinned if False else "-"), flush=True)
    for rep in range(2):
        t = [time.perf_counter()]
        eng = pymf_b200.Engine(d, n, k, device=0); t.append(time.perf_counter())
        eng.upload_x(X); t.append(time.perf_counter())
        eng.set_w(W0); eng.set_h(H0); t.append(time.perf_counter())
        eng.run(20); t.append(time.perf_counter())
        eng.get_w(); eng.get_h(); t.append(time.perf_counter())
        eng.close(); t.append(time.perf_counter())
        names = ["create", "upload_x", "set_w/h", "run20", "get_w/h", "close"]
        print("rep %d: " % rep + "  ".join("%s %.3f" % (nm, t[i + 1] - t[i]) for i, nm in enumerate(names)),
              " upload GB/s %.1f" % (X.nbytes / 1e9 / (t[2] - t[1])), flush=True)
    # raw host memcpy speed for reference
    Y = np.empty_like(X)
    t0 = time.perf_counter(); np.copyto(Y, X); t1 = time.perf_counter()
    print("numpy copy 4 GiB single thread: %.3f s = %.1f GB/s" % (t1 - t0, X.nbytes / 1e9 / (t1 - t0)))
    return
    import torch
    Xt = torch.from_numpy(X)
    t0 = time.perf_counter(); Xp = Xt.pin_memory(); t1 = time.perf_counter()
    print("torch pin_memory (alloc+copy): %.3f s" % (t1 - t0))
    torch.cuda.synchronize()
    Xd = torch.empty((d, n), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    t0 = time.perf_counter(); Xd.copy_(Xp, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("pinned H2D 4 GiB: %.3f s = %.1f GB/s" % (t1 - t0, X.nbytes / 1e9 / (t1 - t0)))
    t0 = time.perf_counter(); Xd.copy_(Xt); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("pageable H2D 4 GiB (torch): %.3f s = %.1f GB/s" % (t1 - t0, X.nbytes / 1e9 / (t1 - t0)))
    print("cpus", os.cpu_count(), len(os.sched_getaffinity(0)))


if __name__ == "__main__":
    main()
