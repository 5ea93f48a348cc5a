"""Experiment helper: per-stage timeline of the fused kernel (PYMFB_TRACE build): who waits for whom."""
import sys
import numpy as np

a = np.fromfile(sys.argv[1], dtype=np.int64).reshape(-1, 16)
lo, hi = 400, 3600
typ = a[lo:hi, 13]


def span(x, y, label, sel=None):
    m = (a[lo:hi, x] > 0) & (a[lo:hi, y] > 0)
    if sel is not None:
        m &= (typ == sel)
    d = (a[lo:hi, y] - a[lo:hi, x])[m]
    if len(d):
        print("%-56s n %5d mean %7.0f  median %7.0f  p90 %7.0f" % (label, len(d), d.mean(), np.median(d), np.percentile(d, 90)))


for sel, nm in ((1, "A stages"), (2, "B stages")):
    print("==", nm)
    span(0, 1, "prod: wait for empty slot", sel)
    span(1, 3, "TMA issue -> conv sees data (load latency + conv queue)", sel)
    span(2, 9, "conv: wait A slot (aempty)", sel)
    span(9, 3, "conv: wait data (full)", sel)
    span(3, 4, "conv: work", sel)
    span(4, 6, "conv done -> mma past wait", sel)
    span(5, 6, "mma: wait afull", sel)
    span(6, 7, "mma: issue + commit", sel)
    span(1, 7, "TMA issue -> mma issued (stage lifetime)", sel)
for slot, nm in ((1, "producer issue"), (4, "convert done"), (7, "mma issued")):
    v = a[lo:hi, slot]; v = np.sort(v[v > 0]); d = np.diff(v)
    print("%-18s interval mean %.0f median %.0f cycles" % (nm, d.mean(), np.median(d)))
