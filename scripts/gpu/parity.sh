#!/bin/bash
mkdir -p gpurun_out
(echo "== v2"; timeout 300 python tools/parity_report.py fp16 128 2>&1 | grep -v "^CVAE"; echo "== v1"; BP_ENGINE_V1=1 timeout 300 python tools/parity_report.py fp16 128 2>&1 | grep -v "^CVAE"; echo "== v2 512";  timeout 300 python tools/parity_report.py fp16 512 2>&1 | grep -v "^CVAE") > gpurun_out/parity.log 2>&1
cat gpurun_out/parity.log
for c in 32 64; do BP_CHUNK=$c timeout 600 python bench.py --precision fp16 --tiles 256 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v2_c$c.json 2>/dev/null; cut -c1-200 gpurun_out/bench_v2_c$c.json | tail -1; done
