"""Summarise an `ncu --metrics a,b,c --csv` log (one row per launch and metric) per kernel name: mean of every metric.
usage: python tools/ncu_metrics_summary.py log.csv > summary.txt"""
import csv
import re
import sys
from collections import defaultdict

rows = [l for l in open(sys.argv[1], newline="") if l.startswith('"')]
acc = defaultdict(lambda: defaultdict(list))
units = {}
for r in csv.DictReader(rows):
    name = re.sub(r"\(.*$", "", re.sub(r"^void ", "", r["Kernel Name"]))
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    acc[name][r["Metric Name"]].append(v)
    units[r["Metric Name"]] = r["Metric Unit"]
tot = {k: sum(v.get("gpu__time_duration.sum", [0])) for k, v in acc.items()}
print(f"# {sys.argv[1]}: per-kernel means over the captured launches (warm caches: --cache-control none --clock-control none)")
for k in sorted(acc, key=lambda k: -tot[k]):
    n = max(len(v) for v in acc[k].values())
    print(f"---- {k}  ({n} launches)")
    for m, vals in sorted(acc[k].items()):
        print(f"   {m:78s} {sum(vals) / len(vals):14.4f} {units[m]}")
