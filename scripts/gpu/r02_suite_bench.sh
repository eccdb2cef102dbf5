#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest4.log 2>&1
echo "pytest exit $?"; grep -E "^E +assert [0-9.e-]+ <=|AssertionError|NotImplemented|Error|passed|failed" gpurun_out/r02_pytest4.log | cut -c1-300 | head -20
timeout 1200 python bench.py --profile-layers > gpurun_out/r02_bench4.json 2> gpurun_out/r02_bench4.err
echo "bench exit $?"; python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r02_bench4.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], "parity", d["parity"]["max_rel_l2"])
    for k in ("fp32", "cgan", "variance", "lightcone"):
        v = d.get(k)
        print(k, {a: b for a, b in v.items() if a in ("value", "unit", "ms_per_step", "parity", "stages_s_max_over_ranks")} if v else v)
    print("cpu", d["cpu_baseline"]["value"], "roof", d["roofline"]["frac"], d["roofline"]["frac_burst"], d["roofline"]["whole_net"])
except Exception as e:
    print("parse failed", e)
PY
tail -6 gpurun_out/r02_bench4.err | cut -c1-200
