#!/bin/bash
# full capture of four spline-sweep launches of a line of sight (largest crops come first)
mkdir -p gpurun_out
CMD="python bench.py --config lightcone --steps 1 --warmup 1"
timeout 600 ncu --set full --clock-control none -k regex:zoom_sweep -s 8 -c 4 -o gpurun_out/prof_zoom_r02 $CMD > gpurun_out/ncu_zoom.log 2>&1
ncu -i gpurun_out/prof_zoom_r02.ncu-rep --page raw --csv > gpurun_out/r02_zoom_raw.csv 2>/dev/null
rm -f gpurun_out/prof_zoom_r02.ncu-rep
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02_zoom_raw.csv')))
h=rows[0]
cols=['Kernel Name','Grid Size','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__cycles_active.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__shared_mem_per_block_dynamic','smsp__issue_active.avg.pct_of_peak_sustained_active']
idx=[h.index(c) for c in cols if c in h]
w=csv.writer(open('gpurun_out/r02_ncu_zoom_full.csv','w'))
w.writerow([h[i] for i in idx]); w.writerow([rows[1][i] for i in idx])
for r in rows[2:]:
    w.writerow([r[i] for i in idx]); print([r[i][:40] for i in idx])
PY
