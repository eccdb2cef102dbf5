#!/bin/bash
# BASELINE config 5: synthetic lightcone (planes of 4096^2, 12 redshift slices) on 1 and N GPUs
N=${1:-2}
ARGS="--synthetic --n-plane 12 --output-resolution 1549 --synthetic-plane-pixels 4096 --batch 64"
timeout 900 python scripts/create_lightcone.py $ARGS --output-file gpurun_out/y1 2>&1 | grep -E "Painted|Projected|Using [0-9]+ tiles" | tail -14
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/create_lightcone.py $ARGS --output-file gpurun_out/yN 2>&1 | grep -E "Painted|Projected" | tail -3
rm -f gpurun_out/y1*.npy gpurun_out/yN*.npy
