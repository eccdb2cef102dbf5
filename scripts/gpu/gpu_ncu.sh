#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --precision fp16 --tiles 16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_small.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_win -s 1 -c 2 -o gpurun_out/prof_win_res python bench.py --precision fp16 --tiles 16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/ncu_full.log
ls -la gpurun_out/
