#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py --precision fp16 --tiles 256 --steps 3 --warmup 3 --cpu-tiles 4 --profile-layers > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err
cat gpurun_out/bench_fp16.json; tail -27 gpurun_out/bench_fp16.err
# ncu launch list of a short bench run (after the plain run above exited 0)
timeout 600 python bench.py --precision fp16 --tiles 16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_small.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_v1.csv python bench.py --precision fp16 --tiles 16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/ncu_launch.log
