"""Summarise an `ncu --page source --csv` dump (SASS view): stall samples and executed instructions by region.
usage: python tools/ncu_src_summary.py file_src.csv [bucket]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))[2:]
bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 100
tot = sum(int(r[2]) for r in rows)
toti = sum(int(r[5]) for r in rows)
print("samples", tot, "warp-instructions", toti)
for b in range(0, len(rows), bucket):
    sl = rows[b:b + bucket]
    s = sum(int(r[2]) for r in sl)
    n = sum(int(r[5]) for r in sl)
    ops = []
    for r in sl:
        t = r[1].split()
        ops.append(t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else ""))
    c = collections.Counter(ops).most_common(5)
    print(f"{b:5d} samples {s:7d} {100 * s / max(tot, 1):5.1f}%  inst {n:10d}  {c}")
