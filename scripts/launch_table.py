#!/usr/bin/env python
"""Per-launch table from an ncu --csv log with several metrics per launch (time, tensor pipe %, DRAM bytes):
one line per launch: id, kernel, grid, us, tensor %, DRAM MB, DRAM GB/s, and (when captured) the bytes TMA loads pulled through the
L2 -> SM port with their rate and the L2 slice utilisation.   usage: launch_table.py launches.csv"""
import csv
import re
import sys
from collections import OrderedDict

rows = OrderedDict()
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    d = rows.setdefault(int(r["ID"]), {"name": re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", ""), "grid": r["Grid Size"]})
    d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
L2 = "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum"
LTS = "lts__t_sectors_srcunit_tex.avg.pct_of_peak_sustained_elapsed"
print("%4s %-44s %-14s %9s %8s %9s %8s %10s %9s %6s" % ("id", "kernel", "grid", "us", "tensor%", "dram_MB", "GB/s", "tma_ld_MB", "tma_GB/s", "lts%"))
for i, d in rows.items():
    us = d.get("gpu__time_duration.sum", 0.0) / 1e3
    mb = (d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)) / 1e6
    l2 = d.get(L2, 0.0) / 1e6
    print("%4d %-44s %-14s %9.1f %8.1f %9.1f %8.0f %10.1f %9.0f %6.1f" % (i, d["name"][:44], d["grid"], us,
          d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0), mb, mb / us * 1e3 if us else 0,
          l2, l2 / us * 1e3 if us else 0, d.get(LTS, 0.0)))
