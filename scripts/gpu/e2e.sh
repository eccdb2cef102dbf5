#!/bin/bash
for st in 32 64 128; do echo "== host step $st"; BP_HOST_STEP=$st timeout 600 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e'])"; done
