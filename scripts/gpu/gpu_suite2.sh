#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/layer_report.py fp16 2>&1 | grep -v "^CVAE" > gpurun_out/layers.log; cat gpurun_out/layers.log | head -3
timeout 600 python bench.py --precision fp16 --tiles 256 --steps 3 --warmup 3 --no-cpu-baseline --profile-layers > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err
cut -c1-400 gpurun_out/bench_fp16.json; tail -27 gpurun_out/bench_fp16.err
