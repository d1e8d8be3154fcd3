#!/usr/bin/env python
"""Summarise an `ncu --csv --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]` log:
per kernel name totals (time, share, DRAM bytes).  usage: summarize_launches.py log.csv [out.json]"""
import csv, collections, json, sys
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
tot_us = tot_b = 0.0
per_launch = collections.defaultdict(dict)
for r in csv.DictReader(lines):
    name = r["Kernel Name"].split("(")[0].replace("void ", "")[:60]
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "")
    m = r["Metric Name"]
    if m == "gpu__time_duration.sum":
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        agg[name][0] += 1; agg[name][1] += us; tot_us += us
    elif m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        agg[name][2] += v * mult; tot_b += v * mult
n = sum(v[0] for v in agg.values())
print(f"# total {tot_us/1e3:.3f} ms over {n} launches; DRAM traffic {tot_b/1e9:.3f} GB")
print("kernel,launches,total_us,share_pct,avg_us,dram_MB")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k},{v[0]},{v[1]:.1f},{100*v[1]/tot_us:.2f},{v[1]/v[0]:.1f},{v[2]/1e6:.1f}")
if len(sys.argv) > 2:
    json.dump({"unit": "one eager U-Net step, 16f 320x512 (ncu --cache-control none --clock-control none)",
               "launches": n, "sum_kernel_ms": tot_us / 1e3, "dram_bytes": tot_b}, open(sys.argv[2], "w"))
