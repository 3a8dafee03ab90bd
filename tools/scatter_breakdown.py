"""Per-kernel table (time, DRAM bytes) of the last iteration in an ncu launch list of `bench.py --scatter-only`.
usage: python tools/scatter_breakdown.py gpurun_out/scatter_launches.csv"""
import csv
import re
import sys
from collections import OrderedDict

rows = [l for l in open(sys.argv[1], newline='') if l.startswith('"')]
L = OrderedDict()
for r in csv.DictReader(rows):
    k = r["ID"]
    n = re.sub(r"\(.*$", "", re.sub(r"^void ", "", r["Kernel Name"]))[:60]
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    L.setdefault(k, {'name': n, 'grid': r['Grid Size']})
    m = r["Metric Name"]
    if m == "gpu__time_duration.sum":
        L[k]['t'] = v * {'ns': 1e-3, 'us': 1, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1, 'msecond': 1e3}[u]
    else:
        L[k][m] = v * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1, 'Gbyte': 1e3}[u]
ids = list(L)
last = [i for i, k in enumerate(ids) if 'k_mark_points' in L[k]['name']][-1]
tot = 0
print("# kernel, us, DRAM read MB, DRAM write MB, DRAM GB/s, grid  (last iteration; ncu per-launch times are cold-cache and serialised)")
for k in ids[last:]:
    e = L[k]
    if not e['name'].startswith('dfb::'):
        continue
    rd = e.get('dram__bytes_read.sum', 0); wr = e.get('dram__bytes_write.sum', 0)
    tot += e['t']
    print('%-34s %8.1f us  rd %7.1f MB  wr %7.1f MB  %6.0f GB/s  grid %s' % (e['name'], e['t'], rd, wr, (rd + wr) / e['t'] * 1e3, e['grid']))
print("# sum of dfb kernels: %.1f us" % tot)
