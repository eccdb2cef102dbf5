#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py --precision fp16 --tiles 256 --steps 3 --warmup 3 --cpu-tiles 4 --profile-layers > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err
cat gpurun_out/bench_fp16.json; tail -32 gpurun_out/bench_fp16.err
