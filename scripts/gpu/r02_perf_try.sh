#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_cvae.py tests/test_gpu_cgan.py -m gpu -q -x > gpurun_out/r02_pytest5.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/r02_pytest5.log | cut -c1-200
for cfg in "64 16" "128 16" "96 32" "128 32"; do
  set -- $cfg
  BP_HOST_STEP=$1 BP_HOST_EDGE=$2 timeout 600 python bench.py --no-extra --no-fp32 --no-cpu-baseline --no-parity --profile-layers > gpurun_out/r02_bench6_$1_$2.json 2> gpurun_out/r02_bench6_$1_$2.err
  echo "step $1 edge $2:" $(grep -o '"value": [0-9.]*' gpurun_out/r02_bench6_$1_$2.json | head -2)
done
grep "p_y_z_in\|p_mu_out" gpurun_out/r02_bench6_64_16.err | sed "s/.*tensor=[0-9] *//" | awk '{printf "%s ", $1}'; echo
