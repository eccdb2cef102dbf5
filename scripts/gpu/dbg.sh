#!/bin/bash
mkdir -p gpurun_out
for d in 0 15; do
echo "== dbg $d"
BP_V2_DBG=$d BP_WIN_TIMING=1 timeout 300 python bench.py --precision fp16 --tiles 64 --steps 1 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "^\[wconv\]" | tail -16 | awk 'NR==1||NR==2||NR==3||NR==4||NR==12||NR==13||NR==14||NR==15||NR==16{print}' | cut -c1-350
done > gpurun_out/dbg.log 2>&1
cat gpurun_out/dbg.log
timeout 300 python tools/layer_report.py fp16 2>&1 | grep -v "^CVAE" | head -3
