"""Per-kernel counts of the Blackwell-native SASS mnemonics in the built library (cuobjdump -sass): UTCHMMA (tcgen05.mma),
UTMALDG / UTMASTG / UBLKCP (TMA), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit), HMMA (legacy mma.sync: must be 0).
usage: python tools/sass_summary.py [lib.so] > profiles/r02_sass_summary.txt"""
import os
import re
import subprocess
import sys
from collections import Counter, defaultdict

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "deflow_b200", "lib", "libdeflow_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
MN = ["UTCHMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "HMMA", "LDGSTS", "ATOMG", "REDG", "RED"]
per = defaultdict(Counter)
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        per[cur]["_total"] += 1
        if op in MN:
            per[cur][op] += 1
dem = subprocess.run(["cu++filt"] + list(per), capture_output=True, text=True).stdout.splitlines() if per else []
names = dict(zip(per, dem)) if len(dem) == len(per) else {k: k for k in per}
print(f"# cuobjdump -sass {os.path.relpath(lib)} (sm_100a): SASS mnemonic counts per kernel")
print(f"# {'kernel':70s} {'instrs':>7s} " + " ".join(f"{m:>7s}" for m in MN))
tot = Counter()
for k in sorted(per, key=lambda k: -per[k]["UTCHMMA"] * 100000 - per[k]["_total"]):
    c = per[k]
    nm = re.sub(r"\(.*$", "", names[k].replace("(int)", "").replace("(bool)", "")).replace("void ", "").replace("dfb::", "")
    print(f"  {nm[:70]:70s} {c['_total']:7d} " + " ".join(f"{c[m]:7d}" for m in MN))
    tot.update(c)
print(f"  {'TOTAL':70s} {tot['_total']:7d} " + " ".join(f"{tot[m]:7d}" for m in MN))
