"""Top stall-sample instructions from an `ncu --page source --csv` export (gz ok): python scripts/ncu_hot.py file [kernel_index] [min_pct]"""
import csv, gzip, io, sys
f = sys.argv[1]
ki = int(sys.argv[2]) if len(sys.argv) > 2 else 0
minp = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
fh = io.TextIOWrapper(gzip.open(f)) if f.endswith(".gz") else open(f)
rows = list(csv.reader(fh))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
s = starts[ki]
e = starts[ki + 1] if ki + 1 < len(starts) else len(rows)
print(rows[s][1][:100])
body = rows[s + 2:e]
tot = sum(int(r[2]) for r in body if r[2].isdigit())
print("total samples", tot)
for i, r in enumerate(body):
    if r[2].isdigit() and int(r[2]) > tot * minp / 100:
        print(f"{i:5d} {r[1].strip()[:84]:84s} {int(r[2]):7d} {100 * int(r[2]) / tot:5.1f}%  exec={r[5]}")
