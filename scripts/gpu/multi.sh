#!/bin/bash
# two-GPU checks: weak-scaling bench line and the sharded lightcone driver (NCCL reduce)
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_g$N.json 2> gpurun_out/bench_g$N.err; tail -2 gpurun_out/bench_g$N.err; cat gpurun_out/bench_g$N.json | cut -c1-700
timeout 900 python scripts/create_lightcone.py --synthetic --n-plane 5 --output-file gpurun_out/y_map_1gpu --output-resolution 256 --synthetic-plane-pixels 1024 > gpurun_out/lc1.log 2>&1; tail -2 gpurun_out/lc1.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/create_lightcone.py --synthetic --n-plane 5 --output-file gpurun_out/y_map_${N}gpu --output-resolution 256 --synthetic-plane-pixels 1024 > gpurun_out/lc$N.log 2>&1; tail -2 gpurun_out/lc$N.log
python - <<PY
import numpy as np
a=np.load("gpurun_out/y_map_1gpu.npy"); b=np.load("gpurun_out/y_map_${N}gpu.npy")
print("y-map 1 GPU vs $N GPUs (different device RNG draws per tile -> statistical agreement only): shapes", a.shape, b.shape, "finite", np.isfinite(a).all(), np.isfinite(b).all(), "rel diff of sums", abs(a.sum()-b.sum())/abs(a.sum()))
PY
rm -f gpurun_out/y_map_*.npy
