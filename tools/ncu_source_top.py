"""Top stall locations per kernel from an `ncu --page source --print-source sass --csv` export."""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
kern, header, rows = None, None, []


def flush():
    if not kern or not rows:
        return
    tot = sum(int(r["# Samples"] or 0) for r in rows)
    print(f"==== {kern}   total samples {tot}")
    stalls = [c for c in header if c.startswith("stall_") and "Not Issued" not in c]
    agg = {c: sum(int(r[c] or 0) for r in rows) for c in stalls}
    print("   by reason:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tot * 0.01})
    for i, r in enumerate(rows):
        r["_i"] = i
    for r in sorted(rows, key=lambda r: -int(r["# Samples"] or 0))[:top]:
        n = int(r["# Samples"] or 0)
        why = {c[6:]: int(r[c] or 0) for c in stalls if int(r[c] or 0) > n * 0.15}
        print(f"   {n:7d} {100.0 * n / max(tot, 1):5.1f}%  #{r['_i']:5d} ex={r['Instructions Executed']:>8}  {r['Source'].strip()[:70]:70s} {why}")


with open(path) as f:
    for rec in csv.reader(f):
        if not rec:
            continue
        if rec[0] == "Kernel Name":
            flush()
            kern, header, rows = rec[1][:110], None, []
        elif rec[0] == "Address":
            header = rec
        elif header:
            rows.append(dict(zip(header, rec)))
flush()
