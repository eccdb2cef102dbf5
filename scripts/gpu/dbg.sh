#!/bin/bash
# which resource bounds each layer: rerun the per-CTA cycle counters with parts of the kernel disabled
# (BP_V2_DBG bits: 1 no TMEM loads, 2 no stores, 4 no weight copies, 8 no patch copies)
for d in 0 8 4 12 2 3 15; do
  echo "== BP_V2_DBG=$d"
  BP_V2_NOTUNE=1 BP_V2_DBG=$d BP_WIN_TIMING=1 timeout 300 python bench.py --precision fp16 --tiles 64 --steps 1 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "^\[wconv\]" | tail -17 | awk '{printf "%s %s %s %s %s | ", $2,$3,$13,$14,$15} END{print ""}'
done
