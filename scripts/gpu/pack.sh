#!/bin/bash
# per-layer formulation sweep (BP_V2_PACK = cin:cout:k:G:Jy)
for pk in 16:8:7:1:8 16:8:7:1:4 16:8:7:2:4 16:8:7:2:8 16:8:7:1:16 8:1:5:4:8 8:1:5:4:16 8:1:5:2:16 8:1:5:4:4; do
  for nr in 0 1; do
    if [ $nr = 1 ]; then export BP_V2_NORESIDENT=1; else unset BP_V2_NORESIDENT; fi
    echo -n "pack $pk noresident=$nr: "
    BP_V2_PACK=$pk timeout 300 python bench.py --precision fp16 --tiles 256 --steps 2 --warmup 3 --no-cpu-baseline --profile-layers 2>&1 >/dev/null | grep -E "p_mu_out        [01]" | awk '{printf "%s.%s %s ms | ", $1,$2,$9}'; echo
  done
done
