"""GPU experiment (not a test): the CTA-pair H-update kernel (PYMFB_TS2=1, kernels_ts2.cuh) against the default TS
kernel on a kp = 64 shape with an odd number of column tiles, then H-only timing on 8192 x 524288."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pymf_b200  # noqa: E402


def one(d, n, k, ts2, iters):
    if ts2:
        os.environ["PYMFB_TS2"] = "1"
    else:
        os.environ.pop("PYMFB_TS2", None)
    e = pymf_b200.Engine(d, n, k, path="tc")
    e.gen_x(1); e.gen_w(2); e.gen_h(3)
    e.run(1, compute_w=False, compute_h=True, compute_err=False, early_stop=False)
    H = e.get_h(np.float32) if d * n < 2 ** 26 else None
    e.sync()
    ev0, ev1 = e.event(), e.event()
    e.record(ev0)
    e.enqueue(iters, compute_w=False, compute_h=True, compute_err=False)
    e.record(ev1)
    e.sync()
    ms = e.elapsed_ms(ev0, ev1) / iters
    f, _ = e.run(2, early_stop=False)
    e.close()
    return H, ms, f


if __name__ == "__main__":
    for shape in ((1024, 128 * 33, 64),):
        Ha, ta, fa = one(*shape, ts2=False, iters=3)
        Hb, tb, fb = one(*shape, ts2=True, iters=3)
        print(shape, "bit-identical:", np.array_equal(Ha, Hb), "max abs diff", float(np.max(np.abs(Ha - Hb))),
              "ferr", fa, fb, flush=True)
    for ts2 in (False, True, False, True):
        _, ms, _ = one(4096, 32768, 32, ts2, 20)
        stages = 129.0 * 256 / float(os.environ.get("PYMFB_GRID", "148"))
        print("4096x32768 k=32 H-only, grid %s  ts2=%s: %.4f ms per pass = %.0f ns per 16 KB stage per CTA" % (
            os.environ.get("PYMFB_GRID", "148"), ts2, ms, ms * 1e6 / stages), flush=True)
