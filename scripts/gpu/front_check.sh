#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cvae.py -m gpu -x -q -k "fp16" 2>&1 | tail -3
timeout 600 python bench.py --precision fp16 --tiles 256 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-200
bash scripts/gpu/ncu_front.sh > /dev/null 2>&1
ncu -i gpurun_out/prof_front_r01.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__cycles_active.avg.pct_of_peak_sustained_elapsed 2>/dev/null | cut -d, -f5,12- | cut -c1-200
