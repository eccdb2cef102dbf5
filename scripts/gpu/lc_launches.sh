#!/bin/bash
# per-kernel time of one full line of sight (ncu launch list of the first, warm-up, line of sight)
mkdir -p gpurun_out
CMD="python bench.py --config lightcone --steps 1 --warmup 1"
timeout 600 $CMD > gpurun_out/lc_plain.log 2>&1 || { tail -5 gpurun_out/lc_plain.log; exit 1; }
grep -o '"ms_per_step": [0-9.]*' gpurun_out/lc_plain.log | head -1
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 2200 --csv --log-file gpurun_out/r02_lc_launches.csv $CMD > gpurun_out/lc_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02_lc_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
tot=collections.Counter(); cnt=collections.Counter()
for r in rows[1:]:
    name=r[ki].split('(')[0][:64]
    tot[name]+=float(r[vi].replace(',','')); cnt[name]+=1
for k,v in tot.most_common(22): print(f"{v/1e6:9.3f} ms  x{cnt[k]:5d}  {k}")
print("sum", sum(tot.values())/1e6, "launches", len(rows)-1)
PY
