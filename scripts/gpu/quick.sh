#!/bin/bash
# parity (fp16 CVAE tests) + device / end-to-end throughput
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cvae.py tests/test_gpu_cgan.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do
timeout 300 python bench.py --precision fp16 --tiles 256 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.0f  e2e %.0f tiles/s  ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done
