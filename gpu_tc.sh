#!/bin/bash
mkdir -p gpurun_out
echo "=== fp16 ===" > gpurun_out/layers.log
timeout 120 python tools/layer_report.py fp16 >> gpurun_out/layers.log 2>&1; echo "rc=$?" >> gpurun_out/layers.log
echo "=== fp16 swapped LBO/SBO ===" >> gpurun_out/layers.log
BP_TC_SWAP_LBO_SBO=1 timeout 120 python tools/layer_report.py fp16 >> gpurun_out/layers.log 2>&1; echo "rc=$?" >> gpurun_out/layers.log
echo "=== bf16 ===" >> gpurun_out/layers.log
timeout 120 python tools/layer_report.py bf16 >> gpurun_out/layers.log 2>&1; echo "rc=$?" >> gpurun_out/layers.log
cat gpurun_out/layers.log
