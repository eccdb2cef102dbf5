#!/bin/bash
# ncu evidence for profiles/: (1) launch list of a short bench run, (2) full capture of the window-GEMM kernels
mkdir -p gpurun_out
CMD="python bench.py --precision fp16 --tiles 64 --steps 1 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain_small.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_v2.csv $CMD > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
timeout 300 $CMD > gpurun_out/plain_small2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wconv_kernel -s 64 -c 32 -o gpurun_out/prof_wconv_r01 $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
