#!/bin/bash
# short bench with the per-layer table: $1 = output tag
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python bench.py --no-extra --no-fp32 --no-cpu-baseline --profile-layers > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"; python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], "parity", d["parity"]["max_rel_l2"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("parse failed", e)
PY
grep -E "p_y_z_in|p_mu_out|prior_network" gpurun_out/bench_$TAG.err | awk '{printf "%s%s:%s ", $1,$2,$9} END {print ""}'
tail -3 gpurun_out/bench_$TAG.err | cut -c1-300
