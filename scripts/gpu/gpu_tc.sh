#!/bin/bash
mkdir -p gpurun_out
echo "=== fp16 window ===" > gpurun_out/layers.log
timeout 120 python tools/layer_report.py fp16 >> gpurun_out/layers.log 2>&1; echo "rc=$?" >> gpurun_out/layers.log
cat gpurun_out/layers.log
