#!/bin/bash
# ncu evidence for profiles/: (1) launch list of a short bench run, (2) full capture of the window-GEMM kernels.
# Net creation times the candidate formulations of the packed layers (extra wconv launches at the start of the
# list, counted by BP_V2_TUNE_LOG); the full capture skips those and three warm-up steps.
mkdir -p gpurun_out
CMD="python bench.py --precision fp16 --tiles 256 --steps 1 --warmup 3 --no-cpu-baseline"
BP_V2_TUNE_LOG=1 timeout 300 $CMD > gpurun_out/plain_small.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r01_v3.csv $CMD > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
T=$(grep "^\[tune\]" gpurun_out/plain_small.log | tail -1 | sed 's/.*so far \([0-9]*\)).*/\1/')
W=$(grep -c "wconv_kernel" gpurun_out/launches_r01_v3.csv)
echo "tune launches $T, wconv launches in run $W"
S=$(( T + 3 * 20 ))
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:wconv_kernel -s $S -c 20 -o gpurun_out/prof_wconv_r01 $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
