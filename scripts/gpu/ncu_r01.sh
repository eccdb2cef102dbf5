#!/bin/bash
# ncu evidence for profiles/: (1) launch list of a short bench run, (2) full capture of the window-GEMM kernels.
# Net creation times the candidate formulations of the packed layers; a profiler serialises launches and would
# perturb that, so the selection of a plain run (BP_V2_TUNE_LOG) is replayed through BP_V2_CHOICES.
mkdir -p gpurun_out
CMD="python bench.py --precision fp16 --tiles 256 --steps 1 --warmup 3 --no-cpu-baseline"
BP_V2_TUNE_LOG=1 timeout 300 $CMD > gpurun_out/plain_small.log 2>&1 || exit 1
export BP_V2_CHOICES=$(grep "^\[tune\] " gpurun_out/plain_small.log | tail -1 | sed 's/.*BP_V2_CHOICES=//')
echo "BP_V2_CHOICES=$BP_V2_CHOICES"
timeout 300 $CMD > gpurun_out/plain_small2.log 2>&1 || exit 1
grep -o '"value": [0-9.]*' gpurun_out/plain_small.log gpurun_out/plain_small2.log | head -2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r01_v3.csv $CMD > gpurun_out/ncu_launch.log 2>&1
tail -1 gpurun_out/ncu_launch.log | cut -c1-200
# full capture: the 20 window-GEMM launches of the fourth (timed) device step
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:wconv_kernel -s 60 -c 20 -o gpurun_out/prof_wconv_r01 $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
# the report (and its source page) is too large to travel back (64 MiB limit): export what profiles/ keeps
ncu -i gpurun_out/prof_wconv_r01.ncu-rep --page raw --csv > gpurun_out/wconv_raw_r01.csv 2>/dev/null
( cd /tmp && ncu -i $OLDPWD/gpurun_out/prof_wconv_r01.ncu-rep --page source --csv --print-source sass > /tmp/wconv_source_r01.csv 2>/dev/null )
for k in 3 4 7 8 17 18; do    # L0, L1, residual conv (open / close), T14, conv7
  python tools/ncu_stalls.py /tmp/wconv_source_r01.csv 40 $k > gpurun_out/wconv_stalls_r01_k$k.txt 2>&1
done
rm -f gpurun_out/prof_wconv_r01.ncu-rep
du -sh gpurun_out
