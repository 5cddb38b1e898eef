#!/usr/bin/env python
"""Markdown table of the bench lines kept under profiles/ (usage: python profiles/summarize.py [prefix])."""
import glob
import json
import os
import sys

here = os.path.dirname(os.path.abspath(__file__))
pre = sys.argv[1] if len(sys.argv) > 1 else "r2_"
print("| file | workload | GPUs | particles (total) | ms/step | updates/s | update ms (roofline frac) | merge ms | weights / estimate / resample ms | production ms/step | exchange_check |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for f in sorted(glob.glob(os.path.join(here, pre + "bench*.json"))):
    try:
        l = json.loads(open(f).read().strip().split("\n")[-1])
    except Exception:
        continue
    if l.get("impl") == "reference":
        print("| %s | %s | CPU x%d | | %.1f | %.3g | | | | | |" % (os.path.basename(f), l["config"]["workload"], l["cpu_baseline"]["cores"], l["ms_per_step"], l["value"]))
        continue
    c, ph = l["config"], l["phase_ms"]
    ec = l.get("exchange_check")
    print("| %s | %s | %d | %s | %.2f | %.1f G | %.2f (%.3f) | %.2f | %.3f / %.3f / %.3f | %s | %s |" % (
        os.path.basename(f), c["workload"], l["n_gpus"], "{:,}".format(c.get("particles_total", c["particles_per_gpu"] * l.get("n_gpus", 1))),
        l["ms_per_step"], l["value"] / 1e9, ph["update"], l["roofline"]["frac"], ph["merge"], ph["weights"], ph["estimate"], ph["resample"],
        ("%.2f" % l["production"]["ms_per_step"]) if l.get("production") else "",
        ("ok, %d migrated" % ec["migrated_checked"]) if ec and ec.get("ok") else ("" if ec is None else "FAILED")))
