"""Experiment helper: summarise a PYMFB_TRACE dump (see kernels_tc.cuh TRACE_AT) of the H-update TS kernel."""
import sys
import numpy as np

a = np.fromfile(sys.argv[1], dtype=np.int64).reshape(-1, 16)
lo, hi = int(sys.argv[2]) if len(sys.argv) > 2 else 200, int(sys.argv[3]) if len(sys.argv) > 3 else 1000
names = {0: "prod: before empty wait", 1: "prod: after empty wait (issue)", 2: "conv: before full wait",
         9: "conv: after full wait", 3: "conv: after aempty wait", 4: "conv: done (arrive afull)",
         5: "mma: before waits", 6: "mma: after full+afull", 7: "mma: issued+committed", 8: "mma: before tempty wait",
         10: "epi: before tfull wait", 11: "epi: after tfull", 12: "epi: drained"}
t0 = a[lo:hi][a[lo:hi] > 0].min()
print("stages %d..%d" % (lo, hi))
for slot in (1, 4, 7):
    v = a[lo:hi, slot]
    v = v[v > 0]
    d = np.diff(v)
    print("%-34s per-stage interval: mean %.0f  median %.0f  p90 %.0f cycles" % (names[slot], d.mean(), np.median(d), np.percentile(d, 90)))


def span(x, y, label):
    m = (a[lo:hi, x] > 0) & (a[lo:hi, y] > 0)
    d = (a[lo:hi, y] - a[lo:hi, x])[m]
    print("%-52s mean %7.0f  median %7.0f  p90 %7.0f" % (label, d.mean(), np.median(d), np.percentile(d, 90)))


span(0, 1, "prod wait for empty slot")
span(1, 9, "TMA issue -> data seen by convert (load latency)")
span(2, 9, "conv wait for data")
span(9, 3, "conv wait for A slot")
span(3, 4, "conv work (LDS + split + STTM + wait)")
span(4, 6, "conv done -> mma past its waits")
span(5, 6, "mma wait (full + afull)")
span(6, 7, "mma issue + commit")
span(1, 7, "TMA issue -> mma committed (stage lifetime - mma exec)")
m = a[lo:hi, 11] > 0
if m.any():
    span(10, 11, "epi wait for segment")
    span(11, 12, "epi drain")
# stage lifetime: issue of stage i -> issue of stage i + STAGES (slot reuse)
iss = a[:, 1]
for depth in (8,):
    v = iss[lo + depth:hi] - iss[lo:hi - depth]
    v = v[(iss[lo + depth:hi] > 0) & (iss[lo:hi - depth] > 0)]
    print("issue(i+%d) - issue(i): mean %.0f cycles  => %.0f cycles/stage" % (depth, v.mean(), v.mean() / depth))
