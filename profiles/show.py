#!/usr/bin/env python
"""Print the key figures of bench.py JSON lines:  python profiles/show.py file [file ...]"""
import json
import sys

for f in sys.argv[1:]:
    try:
        j = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as e:
        print(f, "no JSON line:", e)
        continue
    r = j.get("roofline") or {}
    c = j.get("clocks") or {}
    print("%s\n  N=%s %s value %.2f it/s  %.4f ms/step  h %.3f x %.3f  bound %s frac %.3f (hbm %.3f tensor %.3f)  clocks %s %s launches %s" % (
        f, j.get("n_gpus"), j.get("scaling"), j["value"], j["ms_per_step"], r.get("h_update_ms", 0), r.get("xht_ms", 0), r.get("bound"),
        r.get("frac", 0), r.get("hbm_frac", 0), r.get("tensor_frac", 0), c.get("sm_mhz"), c.get("reasons"), j.get("gpu_launches")))
    for k in ("e2e", "e2e_pageable"):
        e = j.get(k)
        if e and "seconds_total" in e:
            print("  %-13s %.3f it/s  total %.3f s (upload %.3f, iterate %.3f)%s" % (
                k, e["value"], e["seconds_total"], e.get("seconds_upload") or 0, e.get("seconds_iterations") or 0,
                ("  ratio_to_pinned %.2f" % e["ratio_to_pinned"]) if "ratio_to_pinned" in e else ""))
    if j.get("cpu_baseline"):
        cb = j["cpu_baseline"]
        print("  cpu_baseline  %.5g it/s (%s, %s cores, prefix rate %.4g on %s cols)" % (cb["value"], cb["kind"], cb["cores"], cb.get("measured_prefix_value", 0), cb.get("prefix_columns")))
    if j.get("parity_vs_golden"):
        p = j["parity_vs_golden"]
        print("  parity        W %.2e H %.2e ferr %.2e ok=%s (%s)" % (p["rel_w"], p["rel_h"], p["rel_ferr"], p["ok"], p["case"]))
    s = j.get("secondary")
    if s:
        sr = s["roofline"]
        print("  secondary     %.1f it/s %.4f ms  h %.3f x %.3f  %s frac %.3f  clocks %s" % (s["value"], s["ms_per_step"], sr["h_update_ms"], sr["xht_ms"], sr["bound"], sr["frac"], (s.get("clocks") or {}).get("sm_mhz")))
        for k in ("e2e", "e2e_pageable"):
            if s.get(k):
                print("    %-13s %.2f it/s (upload %.3f s)" % (k, s[k]["value"], s[k].get("seconds_upload") or 0))
