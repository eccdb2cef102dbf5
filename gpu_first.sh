#!/bin/bash
# first GPU contact: fp32 path parity
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests/test_gpu_cvae.py -m gpu -x -q -k "fp32 or api or checkpoint or variance" -s 2>&1 | tail -40 > gpurun_out/pytest_fp32.log
cat gpurun_out/pytest_fp32.log
