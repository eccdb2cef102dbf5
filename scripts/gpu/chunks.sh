#!/bin/bash
mkdir -p gpurun_out
for c in 128 256; do
echo "== chunk $c"
BP_CHUNK=$c timeout 600 python bench.py --precision fp16 --tiles 256 --steps 3 --warmup 3 --no-cpu-baseline --profile-layers 2> gpurun_out/bench_c$c.err | cut -c1-160
grep -v "^CVAE" gpurun_out/bench_c$c.err | tail -27 | awk '{printf "%s %s %s %s %s | ", $1,$2,$5,$8,$9} NR%3==0{print ""}'; echo
done
