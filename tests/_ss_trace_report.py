"""Experiment helper: summarise a PYMFB_TRACE dump of the SS H-update kernel (k_h_update_tc, three rings)."""
import sys
import numpy as np

a = np.fromfile(sys.argv[1], dtype=np.int64).reshape(-1, 16)
lo, hi = int(sys.argv[2]) if len(sys.argv) > 2 else 600, int(sys.argv[3]) if len(sys.argv) > 3 else 3000
print("stages %d..%d" % (lo, hi))
for slot, name in ((1, "X producer issue"), (4, "split done"), (7, "mma issued+committed")):
    v = a[lo:hi, slot]
    v = v[v > 0]
    d = np.diff(v)
    print("%-34s per-stage interval: mean %.0f  median %.0f  p90 %.0f cycles" % (name, d.mean(), np.median(d), np.percentile(d, 90)))


def span(x, y, label):
    m = (a[lo:hi, x] > 0) & (a[lo:hi, y] > 0)
    d = (a[lo:hi, y] - a[lo:hi, x])[m]
    if len(d):
        print("%-52s mean %7.0f  median %7.0f  p90 %7.0f" % (label, d.mean(), np.median(d), np.percentile(d, 90)))


span(0, 1, "X producer wait for empty X slot")
span(10, 11, "B producer wait for empty B slot")
span(1, 9, "X TMA issue -> seen by split (load latency)")
span(2, 9, "split wait for X")
span(9, 3, "split wait for lo slot")
span(3, 4, "split work")
span(4, 6, "split done -> mma past its waits")
span(5, 8, "mma wait fullb")
span(8, 6, "mma wait fullx + readyl")
span(6, 14, "mma issue (8 MMAs)")
span(14, 7, "mma commits (3) + syncwarp")
span(1, 7, "X TMA issue -> mma committed")
span(11, 8, "B TMA issue -> mma saw fullb")
m = (a[lo + 1:hi, 5] > 0) & (a[lo:hi - 1, 7] > 0)
d = (a[lo + 1:hi, 5] - a[lo:hi - 1, 7])[m]
print("%-52s mean %7.0f  median %7.0f  p90 %7.0f" % ("mma: end of stage i -> top of stage i+1", d.mean(), np.median(d), np.percentile(d, 90)))
