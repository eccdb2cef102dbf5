#!/bin/bash
# tune everything (table to gpurun_out), GPU suite with the fresh table, bench
mkdir -p gpurun_out
timeout 1500 python -m baryon_painter_b200.tune --log --out gpurun_out/tuning_table_full.txt > gpurun_out/r02_tune_full.log 2>&1
echo "tune exit $?"; grep -c . gpurun_out/tuning_table_full.txt; grep "split" gpurun_out/tuning_table_full.txt | cut -c1-120
export BARYON_PAINTER_TUNING_TABLE=$PWD/gpurun_out/tuning_table_full.txt
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_all.log 2>&1
echo "pytest exit $?"; grep -E "^E +assert [0-9.e-]+ <=|AssertionError|NotImplemented|Error|passed|failed" gpurun_out/r02_pytest_all.log | cut -c1-300 | head -30
timeout 900 python bench.py --profile-layers > gpurun_out/r02_bench3.json 2> gpurun_out/r02_bench3.err
echo "bench exit $?"; grep -o '"value": [0-9.]*' gpurun_out/r02_bench3.json | head -3; grep -o '"fp32": {.*' gpurun_out/r02_bench3.json | cut -c1-600; tail -5 gpurun_out/r02_bench3.err
