#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cgan.py tests/test_gpu_lightcone.py -m gpu -q 2>&1 | tail -25 | tee gpurun_out/pytest_cgan.log
