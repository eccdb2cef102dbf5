#!/bin/bash
# the driver's N-GPU launch of the default bench (sharded lightcone as a secondary field), then the reference arm
N=${1:-2}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
echo "bench N=$N exit $?"
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r02_bench_n$N.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, "e2e", d["e2e"]["value"])
    print("lightcone", d.get("lightcone"))
    print("variance", d.get("variance"))
except Exception as e:
    print("parse failed", e)
PY
tail -4 gpurun_out/r02_bench_n$N.err | cut -c1-300
