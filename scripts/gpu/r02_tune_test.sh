#!/bin/bash
# tune the CVAE (fp16) with logging, then run the GPU suite and a bench with the fresh table
mkdir -p gpurun_out
timeout 900 python -m baryon_painter_b200.tune --log --formats fp16 --no-cgan --out gpurun_out/tuning_cvae_f16.txt > gpurun_out/r02_tune2.log 2>&1
echo "tune exit $?"; grep "conv:16:8:k7\|conv:8:1:k5\|conv:3:16" gpurun_out/r02_tune2.log | cut -c1-160 | tail -60
cat baryon_painter_b200/tuning_table.txt | grep -v "conv:16:8:k7:s1:512x512:f16\|conv:8:1:k5:s1:512x512:f16\|conv:3:16:k5:s1:512x512:f16" > gpurun_out/tuning_merged.txt
grep "conv:16:8:k7:s1:512x512:f16\|conv:8:1:k5:s1:512x512:f16\|conv:3:16:k5:s1:512x512:f16" gpurun_out/tuning_cvae_f16.txt >> gpurun_out/tuning_merged.txt
export BARYON_PAINTER_TUNING_TABLE=$PWD/gpurun_out/tuning_merged.txt
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest2.log 2>&1
echo "pytest exit $?"; tail -25 gpurun_out/r02_pytest2.log
timeout 600 python bench.py --profile-layers --no-cpu-baseline > gpurun_out/r02_bench2.json 2> gpurun_out/r02_bench2.err
echo "bench exit $?"; grep -o '"value": [0-9.]*' gpurun_out/r02_bench2.json | head -2; grep -o '"parity": {[^}]*}' gpurun_out/r02_bench2.json; tail -8 gpurun_out/r02_bench2.err
