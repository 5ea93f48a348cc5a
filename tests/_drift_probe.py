"""Probe (not a test): how the 200-iteration trajectory of the cfg2 prefix deviates from the reference golden on
each kernel family - total relative error, the best uniform scale factor, and the error left after removing it."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pymf_b200
from oracle import cases

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_prefix_200"
c = cases.CASES[name]
g = np.load(os.path.join(os.path.dirname(__file__), "golden", "traj_%s.npz" % name))
X, W0, H0 = cases.build(name)
for path in ("simt", "tc"):
    e = pymf_b200.Engine(c["d"], c["n"], c["k"], path=path)
    e.set_err_mode("trace")
    e.upload_x(X); e.set_w(W0); e.set_h(H0)
    at = 0
    for stop in c["keep"]:
        f, _ = e.run(stop - at, early_stop=False)
        at = stop
        out = []
        for nm, A in (("W", e.get_w()), ("H", e.get_h())):
            R = g["%s_%d" % (nm, stop)].astype(np.float64)
            ws = c.get("w_stride", 1) if nm == "W" else 1
            A = A[::ws]
            s = np.vdot(A, R) / np.vdot(R, R)
            out.append("%s rel %.2e scale-1 %+.2e resid %.2e" % (nm, np.linalg.norm(A - R) / np.linalg.norm(R), s - 1,
                                                                  np.linalg.norm(A - s * R) / np.linalg.norm(R)))
        print(path, "it", stop, " | ".join(out), "ferr rel %.2e" % (abs(f[-1] - g["ferr"][stop - 1]) / g["ferr"][stop - 1]), flush=True)
    e.close()
