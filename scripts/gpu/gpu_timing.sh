#!/bin/bash
mkdir -p gpurun_out
BP_CHUNK=64 BP_WIN_TIMING=1 timeout 600 python bench.py --precision fp16 --tiles 64 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/timing.json 2> gpurun_out/timing.err
grep "^\[win\]" gpurun_out/timing.err | tail -13
