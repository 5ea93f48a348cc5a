#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the short text summaries kept under profiles/.

  python profiles/summarize.py launches gpurun_out/x_launches.csv "<command>" > profiles/rN_launches_*.txt
  python profiles/summarize.py full     gpurun_out/x_prof.ncu-rep "<command>" > profiles/rN_ncu_*.txt

`launches` reads the CSV log of  ncu --metrics gpu__time_duration.sum --csv --log-file ...
`full` reads a --set full report through  ncu -i REP --page raw --csv  (works without a GPU).
"""
import csv
import subprocess
import sys
from collections import OrderedDict

SETUP = ("k_gen_uniform", "k_xx", "k_cast", "k_zero_init")

FULL_METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__block_size",
    "launch__grid_size",
    "launch__shared_mem_per_block_dynamic",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]


def short(name):
    name = name.split("(")[0]
    for p in ("void ", "pymfb::"):
        name = name.replace(p, "")
    return name.strip()


def launches(path, cmd):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            us = v / 1000.0 if unit == "ns" else (v if unit in ("us", "usecond") else v * 1000.0)
            rows.append((short(r["Kernel Name"]), us))
    agg = OrderedDict()
    for k, us in rows:
        n, t = agg.get(k, (0, 0.0))
        agg[k] = (n + 1, t + us)
    tot = sum(t for _, t in agg.values())
    print("# ncu launch list (cold-cache, serialised: compare SHARES, not absolutes)")
    print("# command: %s" % cmd)
    print("%-44s %6s %12s %12s %7s" % ("kernel", "n", "total_us", "avg_us", "share"))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-44s %6d %12.1f %12.1f %6.1f%%" % (k, n, t, t / n, 100 * t / tot))
    loop = {k: v for k, v in agg.items() if not any(k.startswith(s) or s in k for s in SETUP)}
    ltot = sum(t for _, t in loop.values())
    print("\n# shares within the iteration loop only (setup kernels %s excluded):" % (", ".join(SETUP)))
    for k, (n, t) in sorted(loop.items(), key=lambda kv: -kv[1][1]):
        print("%-44s %6.1f%%" % (k, 100 * t / ltot))


def full(path, cmd):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units = r[0], r[1]
    print("# ncu --set full --clock-control none --import-source on (durations under ncu are replayed/cold;")
    print("# the CUDA-event numbers of bench.py are the ones reported)")
    print("# command: %s" % cmd)
    for row in r[2:]:
        print("\nkernel: %s" % short(row[hdr.index("Kernel Name")]))
        for m in FULL_METRICS:
            if m in hdr:
                i = hdr.index(m)
                print("  %-68s %s %s" % (m, row[i], units[i]))
        try:
            rd = float(row[hdr.index("dram__bytes_read.sum")].replace(",", ""))
            wr = float(row[hdr.index("dram__bytes_write.sum")].replace(",", ""))
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            rd *= scale.get(units[hdr.index("dram__bytes_read.sum")], 1.0)
            wr *= scale.get(units[hdr.index("dram__bytes_write.sum")], 1.0)
            print("  %-68s %.4f GB" % ("traffic = dram read + write per launch", (rd + wr) / 1e9))
        except Exception:
            pass


if __name__ == "__main__":
    mode, path = sys.argv[1], sys.argv[2]
    cmd = sys.argv[3] if len(sys.argv) > 3 else ""
    (launches if mode == "launches" else full)(path, cmd)
