#!/bin/bash
# round 2, first call: produce the formulation table on the device, then parity + bench with it
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_gpu.txt 2>&1
timeout 900 python -m baryon_painter_b200.tune --log --out gpurun_out/tuning_table.txt > gpurun_out/r02_tune.log 2>&1
echo "tune exit $?"; tail -3 gpurun_out/r02_tune.log | cut -c1-200
export BARYON_PAINTER_TUNING_TABLE=$PWD/gpurun_out/tuning_table.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/r02_pytest.log
timeout 600 python bench.py --profile-layers > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
echo "bench exit $?"; cat gpurun_out/r02_bench.json | cut -c1-1500; tail -40 gpurun_out/r02_bench.err
