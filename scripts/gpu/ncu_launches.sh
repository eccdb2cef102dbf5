#!/bin/bash
# launch list only (per-kernel durations of one 256-tile step); $1 = output tag
TAG=${1:-x}
mkdir -p gpurun_out
CMD="python bench.py --precision fp16 --tiles 256 --steps 1 --warmup 3 --no-cpu-baseline --no-fp32 --no-extra --no-parity"
timeout 300 $CMD > gpurun_out/plain_$TAG.log 2>&1 || { tail -5 gpurun_out/plain_$TAG.log; exit 1; }
grep -o '"value": [0-9.]*' gpurun_out/plain_$TAG.log | head -1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 130 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_$TAG.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_$TAG.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
sel=[r for r in rows[1:] if 75<=int(r[ii])<=99]
tot=0
for r in sel:
    v=float(r[vi].replace(',','')); tot+=v
    print(r[ii], r[ki][:62], v)
print('sum', tot)
PY
