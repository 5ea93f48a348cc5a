"""Experiment helper: summarise the fused kernel's event log (PYMFB_TRACE build, kernels_fused.cuh FTRACE)."""
import sys
from collections import defaultdict

import numpy as np

raw = np.fromfile(sys.argv[1], dtype=np.uint64)
n = int(min(raw[0], 87000))
ev = raw[1:1 + 3 * n].reshape(-1, 3).astype(np.int64)
names = {10: "prod A start", 11: "prod B wait begin", 12: "prod HH wait begin", 13: "prod wait end",
         20: "mma A end", 21: "mma B end", 22: "mma HH end", 30: "epi A drained", 31: "epi B seg drained", 32: "epi HH drained",
         33: "epi A handed", 40: "pub got", 41: "pub fenced", 42: "pub counted", 50: "upd start", 51: "upd stored", 52: "upd flag"}
t0 = ev[:, 2].min()
by = defaultdict(dict)          # tile -> code -> first clock
for code, val, clk in ev:
    tile = val // 64 if code == 42 else val
    by[tile].setdefault(code, clk - t0)
tiles = sorted(by)
print("events", n, "tiles seen", len(tiles), "span %.1f us" % ((ev[:, 2].max() - t0) / 1.9e3))


def stat(a, b, label):
    d = [by[t][b] - by[t][a] for t in tiles if a in by[t] and b in by[t]]
    if d:
        d = np.array(d) / 1.9e3
        print("%-46s n %4d  mean %6.2f  median %6.2f  p90 %6.2f us" % (label, len(d), d.mean(), np.median(d), np.percentile(d, 90)))


stat(10, 20, "A: producer start -> MMA done")
stat(20, 30, "A: MMA done -> epilogue drained")
stat(30, 33, "A: drained -> REDs issued + handed")
stat(33, 40, "A: handed -> publisher got it")
stat(40, 41, "publisher fence")
stat(41, 42, "publisher counter atomic")
stat(42, 50, "last arrival -> update start (this CTA)")
stat(50, 51, "update: loads + math + stores issued")
stat(51, 52, "update: fence + flag")
stat(11, 13, "B: producer wait for the tile's flag")
stat(10, 11, "A start -> B wait begin (same tile)")
stat(10, 13, "A start -> B may start (chain incl. pipelining)")
stat(13, 21, "B: flag seen -> MMA done")
# per-tile period
a = np.array([by[t][10] for t in tiles if 10 in by[t]]) / 1.9e3
print("A start period: mean %.2f us" % np.diff(np.sort(a)).mean())
