#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zoom.py tests/test_gpu_lightcone.py -m gpu -q -x > gpurun_out/lc_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^E +assert|Error|passed|failed" gpurun_out/lc_pytest.log | cut -c1-250 | head -12
timeout 600 python bench.py --config lightcone --steps 3 --warmup 1 > gpurun_out/lc_try.json 2> gpurun_out/lc_try.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/lc_try.json').read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["resident_vs_host_map_rel_l2"], d["stages_s_max_over_ranks"], d.get("map_check"))
PY
