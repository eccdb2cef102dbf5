#!/bin/bash
# throughput against tiles per step (plan chunk = batch)
for n in 64 128 256 512; do
  BP_CHUNK=$n timeout 300 python bench.py --precision fp16 --tiles $n --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > /tmp/line.json
  python - "$n" <<'PY'
import json, sys
d = json.load(open("/tmp/line.json"))
print("tiles/step %s: value %.0f  e2e %.0f tiles/s  %.3f ms/step" % (sys.argv[1], d["value"], d["e2e"]["value"], d["ms_per_step"]))
PY
done
