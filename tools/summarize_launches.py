"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name.
usage: python tools/summarize_launches.py launches.csv [n_steps] > summary.txt"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = [l for l in open(path, newline="") if l.startswith('"')]
tot = defaultdict(float)
cnt = defaultdict(int)
for r in csv.DictReader(rows):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    if r["Metric Unit"] in ("us", "usecond"):
        v *= 1e3
    elif r["Metric Unit"] in ("ms", "msecond"):
        v *= 1e6
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
print(f"# {path}: sum of kernel durations per step: {total / steps / 1e6:.3f} ms ({steps:g} steps captured)")
print("# ms/step  share  launches/step  kernel")
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"{tot[k] / steps / 1e6:8.3f}  {100 * tot[k] / total:5.1f}%  {cnt[k] / steps:7.1f}  {k[:150]}")
