#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share.
usage: summarize_launches.py launches.csv [skip_first_n_launches]"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "s": 1e9, "second": 1e9}.get(unit, 1)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    rows.append((int(r["ID"]), name, ns, r.get("Grid Size", ""), r.get("Block Size", "")))
rows = rows[skip:]
tot = sum(r[2] for r in rows)
agg = defaultdict(lambda: [0, 0.0])
for _, n, ns, _, _ in rows:
    agg[n][0] += 1
    agg[n][1] += ns
print("launches %d  total %.3f ms" % (len(rows), tot / 1e6))
print("%-60s %6s %10s %7s %9s" % ("kernel", "count", "total_ms", "share", "avg_us"))
for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-60s %6d %10.3f %6.1f%% %9.1f" % (n[:60], c, ns / 1e6, 100 * ns / tot, ns / c / 1e3))
