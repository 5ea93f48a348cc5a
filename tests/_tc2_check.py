"""GPU experiment (not a test): the CTA-pair SS H-update kernel (PYMFB_TC2=1, kernels_tc2.cuh) against the default SS
kernel on k = 128 shapes with an odd number of column tiles / ragged sizes, then H-only timing on a cfg3 shard."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pymf_b200  # noqa: E402


def one(d, n, k, tc2, iters):
    if tc2:
        os.environ["PYMFB_TC2"] = "1"
    else:
        os.environ.pop("PYMFB_TC2", None)
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
    for shape in ((1024, 128 * 33, 128), (700, 128 * 9 + 5, 100), (2048, 128 * 100 + 17, 256)):
        Ha, ta, fa = one(*shape, tc2=False, iters=3)
        Hb, tb, fb = one(*shape, tc2=True, iters=3)
        print(shape, "bit-identical:", np.array_equal(Ha, Hb), "max abs diff", float(np.max(np.abs(Ha - Hb))),
              "ferr", fa, fb, flush=True)
    if len(sys.argv) > 1:
        d, n, k = 16384, 131072, 128
        for tc2 in (False, True, False, True):
            _, ms, _ = one(d, n, k, tc2, 20)
            print("%dx%d k=%d H-only  tc2=%s: %.4f ms per pass" % (d, n, k, tc2, ms), flush=True)
