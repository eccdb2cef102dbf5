#!/bin/bash
# A/B of two builds of the library on one box: ab/lib_old.so vs ab/lib_new.so, two rounds each (interleaved)
mkdir -p gpurun_out
LIB=baryon_painter_b200/lib/libbaryon_painter_b200.so
for round in 1 2; do
  for v in old new; do
    cp ab/lib_$v.so $LIB
    timeout 600 python bench.py --no-extra --no-fp32 --no-cpu-baseline --no-parity --profile-layers > gpurun_out/ab_${v}_$round.json 2> gpurun_out/ab_${v}_$round.err
    python - <<PY
import json
d = json.loads(open('gpurun_out/ab_${v}_$round.json').read().strip().splitlines()[-1])
print("$v $round", round(d["value"]), round(d["ms_per_step"], 3), d["clocks"]["sm_mhz"])
PY
    grep -E "p_y_z_in|p_mu_out|prior_network" gpurun_out/ab_${v}_$round.err | awk '{printf "%s ", $9} END {print ""}'
  done
done
cp ab/lib_new.so $LIB
