#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / mean / share.

    python profiles/launch_summary.py gpurun_out/launches.csv [name-filter]
"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
flt = sys.argv[2] if len(sys.argv) > 2 else "radet"
d = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = row["Kernel Name"].split("(")[0].replace("void ", "")
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except Exception:
        continue
    u = row["Metric Unit"]
    v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
    d.setdefault(name, []).append(v)
sel = {k: v for k, v in d.items() if flt in k}
tot = sum(sum(v) / len(v) for v in sel.values())
print(f"{'kernel':48s} {'n':>4s} {'mean us':>9s} {'min':>8s} {'max':>8s} {'share':>6s}")
for k, v in sel.items():
    m = sum(v) / len(v)
    print(f"{k[:48]:48s} {len(v):4d} {m:9.2f} {min(v):8.2f} {max(v):8.2f} {100 * m / tot:5.1f}%")
print(f"{'sum of means (one step)':48s} {'':4s} {tot:9.2f}")
