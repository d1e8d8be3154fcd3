#!/usr/bin/env python
"""Summarise an `ncu --csv --metrics gpu__time_duration.sum` log: per kernel name (and grid) totals."""
import csv, collections, sys
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    name = r["Kernel Name"].split("(")[0][:60]
    agg[name][0] += 1; agg[name][1] += us; tot += us
print(f"total {tot/1e3:.3f} ms over {sum(v[0] for v in agg.values())} launches")
print("kernel,launches,total_us,share_pct,avg_us")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k},{v[0]},{v[1]:.1f},{100*v[1]/tot:.2f},{v[1]/v[0]:.1f}")
