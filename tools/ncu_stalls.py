#!/usr/bin/env python
"""Top stalled SASS instructions of one launch in an ncu report:
   ncu -i rep --page source --csv --print-source sass --launch-skip K --launch-count 1 > f.csv; ncu_stalls.py f.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
k = int(sys.argv[3]) if len(sys.argv) > 3 else 0          # which kernel section of the file
print(len(starts) - 1, "kernel sections; showing", k, rows[starts[k]][1][:80])
hdr = rows[starts[k] + 1]
data = [d for d in rows[starts[k] + 2:starts[k + 1]] if len(d) == len(hdr)]
ia = hdr.index("Warp Stall Sampling (All Samples)"); ie = hdr.index("Instructions Executed")
num = lambda s: int(float(s)) if s not in ("", "-") else 0
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(num(d[ia]) for d in data)
print("total samples", tot, "instructions", len(data))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
top = sorted(range(len(data)), key=lambda i: -num(data[i][ia]))[:n]
for i in sorted(top):
    d = data[i]
    st = sorted(((hdr[c][6:], num(d[c])) for c in stall_cols if num(d[c]) > 0), key=lambda x: -x[1])[:3]
    print(i, d[1].strip()[:72].ljust(72), d[ia].rjust(7), d[ie].rjust(9), st)
