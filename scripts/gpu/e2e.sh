#!/bin/bash
# end-to-end (host buffers in and out) throughput for a few pipeline schedules
for cfg in "64 16" "64 64" "64 8" "96 16" "48 16" "128 16"; do
  set -- $cfg
  echo -n "BP_HOST_STEP=$1 BP_HOST_EDGE=$2: "
  BP_HOST_STEP=$1 BP_HOST_EDGE=$2 timeout 300 python bench.py --precision fp16 --tiles 256 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.0f  e2e %.0f tiles/s' % (d['value'], d['e2e']['value']))"
done
