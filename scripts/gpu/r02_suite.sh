#!/bin/bash
# GPU suite + short bench (args: extra pytest args)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q "$@" > gpurun_out/r02_pytest3.log 2>&1
echo "pytest exit $?"; tail -60 gpurun_out/r02_pytest3.log | cut -c1-220
