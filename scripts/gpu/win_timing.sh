#!/bin/bash
# developer build only: per-launch wait breakdown of the window-GEMM roles (BP_WIN_TIMING)
mkdir -p gpurun_out
BP_WIN_TIMING=1 timeout 600 python bench.py --precision fp16 --tiles 256 --steps 1 --warmup 1 --no-cpu-baseline --no-fp32 --no-extra --no-parity > gpurun_out/win_timing.json 2> gpurun_out/win_timing.err
echo "exit $?"
grep "\[wconv\]" gpurun_out/win_timing.err | tail -20 | cut -c1-400
