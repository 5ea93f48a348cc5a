#!/usr/bin/env python
"""Per-instruction stall summary of an ncu --import-source report (works without a GPU):
   python profiles/stalls.py REPORT.ncu-rep KERNEL_REGEX [TOP]
Prints, for the first matching launch, the TOP instructions by stall samples with their two main stall reasons,
and the totals per stall reason."""
import csv
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    sections, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "data": []}
            sections.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and len(r) == len(cur["hdr"]):
            cur["data"].append(r)
    sec = max(sections, key=lambda s: sum(int(r[s["hdr"].index("# Samples")] or 0) for r in s["data"]))
    hdr, data = sec["hdr"], sec["data"]
    isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[isamp] or 0) for r in data)
    print("# %s\n# %d instructions, %d stall samples" % (sec["name"][:100], len(data), tot))
    agg = {}
    for r in data:
        for j in stall:
            agg[hdr[j]] = agg.get(hdr[j], 0) + int(r[j] or 0)
    print("# by reason: " + ", ".join("%s %.1f%%" % (k, 100.0 * v / max(1, sum(agg.values()))) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp] or 0))[:top_n]
    for i in sorted(top):
        r = data[i]
        st = sorted(((hdr[j], int(r[j] or 0)) for j in stall), key=lambda kv: -kv[1])[:2]
        print("%6d %5.1f%%  exec %10s  %-60s %s" % (i, 100.0 * int(r[isamp]) / max(1, tot), r[iex], r[isrc].strip()[:60], st))


if __name__ == "__main__":
    main()
