#!/bin/bash
mkdir -p gpurun_out
timeout 60 ./tools/issue_probe > gpurun_out/issue_probe.log 2>&1; echo "rc=$?" >> gpurun_out/issue_probe.log
cat gpurun_out/issue_probe.log
