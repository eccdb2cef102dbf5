#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_cvae.py -m gpu -q -k "fp32 or api or checkpoint or variance" 2>&1 | tail -30 > gpurun_out/pytest_fp32.log
cat gpurun_out/pytest_fp32.log
python bench.py --precision fp32 --tiles 64 --steps 2 --warmup 3 --cpu-tiles 3 --profile-layers > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err
cat gpurun_out/bench_fp32.json; tail -40 gpurun_out/bench_fp32.err
