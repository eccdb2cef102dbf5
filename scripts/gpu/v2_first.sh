#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/layer_report.py fp16 2>&1 | grep -v "^CVAE" > gpurun_out/layers_v2.log; cat gpurun_out/layers_v2.log
timeout 600 python -m pytest tests/test_gpu_cvae.py -m gpu -x -q -k "fp16" 2>&1 | tail -15 | tee gpurun_out/pytest_v2.log
BP_WIN_TIMING=1 timeout 600 python bench.py --precision fp16 --tiles 64 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/timing_v2.json 2> gpurun_out/timing_v2.err
grep "^\[wconv\]" gpurun_out/timing_v2.err | tail -16
timeout 600 python bench.py --precision fp16 --tiles 256 --steps 3 --warmup 3 --no-cpu-baseline --profile-layers > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err
cut -c1-300 gpurun_out/bench_v2.json; grep -v "^CVAE" gpurun_out/bench_v2.err | tail -27
