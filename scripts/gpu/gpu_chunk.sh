#!/bin/bash
mkdir -p gpurun_out
for c in 16 32 64; do
BP_CHUNK=$c timeout 600 python bench.py --precision fp16 --tiles 256 --steps 3 --warmup 3 --no-cpu-baseline --profile-layers > gpurun_out/bench_c$c.json 2> gpurun_out/bench_c$c.err
echo "chunk $c"; cut -c1-160 gpurun_out/bench_c$c.json; grep -v "^CVAE" gpurun_out/bench_c$c.err | awk '{print $1,$2,$3,$4,$5,$6,$7, $9,$10,$11,$12,$13}' | tail -25
done
