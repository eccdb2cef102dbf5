#!/bin/bash
# front / tail kernel rework: full GPU suite, then a short bench with the per-layer table
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest6.log 2>&1
echo "pytest exit $?"; grep -E "^E +assert|AssertionError|Error|passed|failed" gpurun_out/r02_pytest6.log | cut -c1-300 | head -20
timeout 900 python bench.py --no-extra --no-fp32 --no-cpu-baseline --profile-layers > gpurun_out/r02_bench10.json 2> gpurun_out/r02_bench10.err
echo "bench exit $?"; python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r02_bench10.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], d["e2e"]["sync_call_value"], "parity", d["parity"]["max_rel_l2"], d["clocks"])
except Exception as e:
    print("parse failed", e)
PY
grep -i "front\|tail\|sum\|total" gpurun_out/r02_bench10.err | head -12
