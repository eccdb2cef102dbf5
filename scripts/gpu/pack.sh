#!/bin/bash
# per-layer formulation sweep (BP_V2_PACK = cin:cout:k:G:Jy)
for pk in none 16:8:7:1:4 16:8:7:1:8 8:1:5:4:4 8:1:5:4:8 3:16:5:4:1 3:16:5:4:2; do
    echo -n "pack $pk: "
    BP_V2_PACK=$pk timeout 300 python bench.py --precision fp16 --tiles 256 --steps 2 --warmup 3 --no-cpu-baseline --profile-layers 2>&1 | grep -E "p_mu_out        [01]|p_y_z_in        0 |value" | awk '/value/{match($0,/"value": [0-9.]+/); printf "%s | ", substr($0,RSTART,RLENGTH)} !/value/{printf "%s.%s %s ms | ", $1,$2,$9}'; echo
done
