#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/layer_report.py fp16 2>&1 | grep -v "^CVAE" | head -1
timeout 600 python bench.py --precision fp16 --tiles 256 --steps 3 --warmup 3 --no-cpu-baseline --profile-layers > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err
cut -c1-200 gpurun_out/bench_v2.json; grep -v "^CVAE" gpurun_out/bench_v2.err | tail -27 | awk '{printf "%s %s %s %s %s | ", $1,$2,$5,$8,$9} NR%3==0{print ""}'; echo
