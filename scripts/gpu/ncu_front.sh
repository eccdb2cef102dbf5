#!/bin/bash
# full ncu capture of the elementwise / stencil kernels at both ends of the network
mkdir -p gpurun_out
CMD="python bench.py --precision fp16 --tiles 256 --steps 1 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain_front.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'front_|tail_stencil|sample_z' -s 12 -c 4 -o gpurun_out/prof_front_r01 $CMD > gpurun_out/ncu_front.log 2>&1
tail -3 gpurun_out/ncu_front.log
