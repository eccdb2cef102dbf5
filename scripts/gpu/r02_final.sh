#!/bin/bash
# final check of the round: GPU suite, smoke(), the default bench line (all secondary fields), the reference arm
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_final.log 2>&1
echo "pytest exit $?"; tail -2 gpurun_out/r02_pytest_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
echo "bench exit $?"; python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r02_bench_final.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], d["e2e"]["sync_call_value"], "parity", d["parity"]["max_rel_l2"], d["clocks"])
    for k in ("fp32", "cgan", "variance", "lightcone"):
        v = d.get(k)
        print(k, {a: b for a, b in v.items() if a in ("value", "unit", "ms_per_step", "parity", "stages_s_max_over_ranks", "map_check")} if v else v)
    print("cpu", d["cpu_baseline"], "roof", {k: d["roofline"][k] for k in ("achieved", "frac", "frac_burst", "whole_net", "share_of_step")})
except Exception as e:
    print("parse failed", e)
PY
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
echo "reference arm exit $?"; tail -1 gpurun_out/r02_bench_ref.json | cut -c1-600
