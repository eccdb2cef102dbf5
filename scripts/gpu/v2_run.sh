#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/layer_report.py fp16 2>&1 | grep -v "^CVAE" > gpurun_out/layers_v2.log; head -1 gpurun_out/layers_v2.log; awk '{print $1,$2,$6,$7}' gpurun_out/layers_v2.log | tail -20 | tr '\n' ';'; echo
timeout 600 python -m pytest tests/test_gpu_cvae.py -m gpu -x -q -k "fp16" 2>&1 | tail -5 | tee gpurun_out/pytest_v2.log
BP_WIN_TIMING=1 timeout 600 python bench.py --precision fp16 --tiles 64 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/timing_v2.json 2> gpurun_out/timing_v2.err
grep "^\[wconv\]" gpurun_out/timing_v2.err | tail -16 | awk 'NR<=4||NR>=12' | cut -c1-340
BP_V2_TUNE_LOG=1 timeout 600 python bench.py --precision fp16 --tiles 256 --steps 3 --warmup 3 --no-cpu-baseline --profile-layers > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err
cut -c1-300 gpurun_out/bench_v2.json; grep -v "^CVAE" gpurun_out/bench_v2.err | grep -v "^prior\|^p_z_in" | tail -60
