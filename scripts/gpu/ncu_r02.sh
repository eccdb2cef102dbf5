#!/bin/bash
# Round-2 ncu evidence for profiles/: (1) launch list of one 256-tile step, (2) full capture of the window-GEMM
# kernels of that step (tensor pipe, DRAM bytes, source-page stalls of the top kernels).
# The formulations come from the shipped table (no load-time timing any more), so the profiled run builds the
# same kernels as the plain run.
mkdir -p gpurun_out
CMD="python bench.py --precision fp16 --tiles 256 --steps 1 --warmup 3 --no-cpu-baseline --no-fp32 --no-extra --no-parity"
timeout 300 $CMD > gpurun_out/r02_plain_small.log 2>&1 || { tail -5 gpurun_out/r02_plain_small.log; exit 1; }
grep -o '"value": [0-9.]*' gpurun_out/r02_plain_small.log | head -1
# warm-up 3 + timed 1 + profile 1 steps of 23 launches each, then the e2e steps: capture the first 200 launches
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/r02_ncu_launches.csv $CMD > gpurun_out/r02_ncu_launch.log 2>&1
tail -1 gpurun_out/r02_ncu_launch.log | cut -c1-120
# full capture: the window-GEMM launches of the fourth (timed) device step: 3 warm-up steps x 20 wconv launches are skipped
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:wconv_kernel -s 60 -c 20 -o gpurun_out/prof_wconv_r02 $CMD > gpurun_out/r02_ncu_full.log 2>&1
tail -2 gpurun_out/r02_ncu_full.log | cut -c1-160
ncu -i gpurun_out/prof_wconv_r02.ncu-rep --page raw --csv > gpurun_out/r02_wconv_raw.csv 2>/dev/null
( cd /tmp && ncu -i $OLDPWD/gpurun_out/prof_wconv_r02.ncu-rep --page source --csv --print-source sass > /tmp/wconv_source_r02.csv 2>/dev/null )
for k in 3 4 7 8 16 17 18; do    # L0, L1, residual conv (open / close), T14, conv7, conv5
  python tools/ncu_stalls.py /tmp/wconv_source_r02.csv 40 $((2 * k + 1)) > gpurun_out/r02_wconv_stalls_k$k.txt 2>&1
done
# the other kernels of the step (front / tail): one full capture each
timeout 600 ncu --set full --clock-control none -k regex:"front_|tail_stencil" -s 6 -c 3 -o gpurun_out/prof_misc_r02 $CMD > gpurun_out/r02_ncu_misc.log 2>&1
ncu -i gpurun_out/prof_misc_r02.ncu-rep --page raw --csv > gpurun_out/r02_misc_raw.csv 2>/dev/null
rm -f gpurun_out/prof_wconv_r02.ncu-rep gpurun_out/prof_misc_r02.ncu-rep
du -sh gpurun_out
