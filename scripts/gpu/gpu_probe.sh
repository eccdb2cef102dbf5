#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./tools/umma_probe > gpurun_out/umma_probe.log 2>&1; echo "rc=$?" >> gpurun_out/umma_probe.log
cat gpurun_out/umma_probe.log
