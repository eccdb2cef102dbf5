#!/bin/bash
# lightcone strong scaling on an 8-GPU box: N = 8, 4, 2, 1
mkdir -p gpurun_out
for N in 8 4 2 1; do bash scripts/gpu/r02_lc_multi.sh $N 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$" | cut -c1-400; done
python - <<'PY'
import json
out = []
for n in (1, 2, 4, 8):
    try:
        out.append(json.loads(open('gpurun_out/r02_lc_n%d.json' % n).read().strip().splitlines()[-1]))
    except Exception as e:
        print(n, "missing", e)
json.dump(out, open('gpurun_out/r02_lightcone_scaling.json', 'w'), indent=1)
print([(d["n_gpus"], round(d["ms_per_step"], 1), round(d["e2e"]["ms_per_step"], 1), d.get("map_check", {}).get("sum")) for d in out])
PY
