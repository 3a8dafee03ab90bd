"""DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of every kernel in `ncu --set full` captures,
averaged over the captured launches -> profiles/r01_traffic.json (read by bench.py for `roofline.traffic`).
usage: python tools/traffic_table.py gpurun_out/top_full.ncu-rep [more.ncu-rep ...] > profiles/r01_traffic.json"""
import csv
import io
import json
import re
import subprocess
import sys
from collections import defaultdict

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
acc = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = re.sub(r"\(.*$", "", re.sub(r"^void ", "", r[idx["Kernel Name"]]))
        name = re.sub(r"^(\w+::)+", "", name).replace("(int)", "").replace(", ", ",")
        rd = float(r[idx["dram__bytes_read.sum"]].replace(",", "")) * UNIT[units[idx["dram__bytes_read.sum"]]]
        wr = float(r[idx["dram__bytes_write.sum"]].replace(",", "")) * UNIT[units[idx["dram__bytes_write.sum"]]]
        t = float(r[idx["gpu__time_duration.sum"]].replace(",", "")) * TIME[units[idx["gpu__time_duration.sum"]]]
        a = acc[name]
        a[0] += 1; a[1] += rd; a[2] += wr; a[3] += t
out = {k: {"launches_captured": v[0], "dram_bytes_per_launch": (v[1] + v[2]) / v[0], "dram_read_bytes_per_launch": v[1] / v[0],
           "dram_write_bytes_per_launch": v[2] / v[0], "avg_launch_ms_under_ncu": v[3] / v[0]} for k, v in acc.items()}
json.dump(out, sys.stdout, indent=1, sort_keys=True)
