#!/bin/bash
mkdir -p gpurun_out
for c in 0.478 0; do
BP_RZ_COMP=$c timeout 600 python tools/layer_errors.py fp32 2>&1 | grep -v "^CVAE" | grep -E "painted|p_mu_out|z_mu|p_y_z_in.24|Error|error"
BP_RZ_COMP=$c timeout 1200 python -m pytest tests/test_gpu_cvae.py -m gpu -q -k "fp32 and not ffma" > gpurun_out/r02_fp32_c$c.log 2>&1
echo "=== comp $c: exit $?"; grep -E "^E +assert [0-9.e-]+ <=|AssertionError|NotImplemented|Error|passed|failed" gpurun_out/r02_fp32_c$c.log | cut -c1-300 | head -30
done
