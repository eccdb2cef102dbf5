#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --config lightcone --steps 1 --warmup 1"
timeout 300 $CMD > gpurun_out/lc_plain.log 2>&1 || { tail -5 gpurun_out/lc_plain.log; exit 1; }
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 1719 -c 1719 --csv --log-file gpurun_out/r02_lc_launches.csv $CMD > gpurun_out/lc_ncu.log 2>&1
tail -2 gpurun_out/lc_ncu.log | cut -c1-200
python - <<'PY'
import csv, collections, re
rows = list(csv.reader(l for l in open('gpurun_out/r02_lc_launches.csv') if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
t = collections.defaultdict(float); c = collections.Counter()
for r in rows[1:]:
    name = re.sub(r"<.*", "", r[ki]); v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else v   # -> us
    t[name] += v; c[name] += 1
tot = sum(t.values())
for k, v in sorted(t.items(), key=lambda kv: -kv[1])[:14]:
    print("%-44s %6d launches %10.1f us %5.1f%%" % (k[:44], c[k], v, 100 * v / tot))
print("total %.1f ms" % (tot / 1e3))
PY
