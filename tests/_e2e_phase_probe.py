"""GPU-box experiment (not a test): where does the time of the drop-in call sequence go, call after call?
   python tests/_e2e_phase_probe.py [cfg2|cfg3]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pymf_b200  # noqa: E402

d, n, k = (16384, 1 << 20, 128) if (len(sys.argv) > 1 and sys.argv[1] == "cfg3") else (4096, 262144, 32)
Xh = pymf_b200.pinned_empty((d, n), np.float32)
Xh[:] = 0.5
W0 = pymf_b200.pinned_empty((d, k), np.float64)
H0 = pymf_b200.pinned_empty((k, n), np.float64)
rng = np.random.default_rng(1)
for call in range(5):
    rng.random(out=W0); rng.random(out=H0)
    t0 = time.perf_counter()
    m = pymf_b200.NMF(Xh, num_bases=k)
    t1 = time.perf_counter()
    m.W, m.H = W0, H0
    t2 = time.perf_counter()
    m.factorize(niter=20)
    t3 = time.perf_counter()
    res = (m.W, m.H, m.ferr)
    t4 = time.perf_counter()
    tm = dict(m.timings)
    del m, res
    t5 = time.perf_counter()
    print("call %d: ctor %.3f  set W/H %.3f  factorize %.3f (engine+upload %.3f iterate %.3f download %.3f)  read %.3f  del %.3f  total %.3f" % (
        call, t1 - t0, t2 - t1, t3 - t2, tm["upload_s"], tm["iterate_s"], tm["download_s"], t4 - t3, t5 - t4, t5 - t0), flush=True)
