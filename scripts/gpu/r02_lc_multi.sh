#!/bin/bash
# lightcone strong scaling on N GPUs of one box (arg: N)
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --config lightcone --gpus $N --steps 3 --warmup 1 > gpurun_out/r02_lc_n$N.json 2> gpurun_out/r02_lc_n$N.err
echo "exit $?"; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"stages_s_max_over_ranks": {[^}]*}\|tiles per rank: [^)]*' gpurun_out/r02_lc_n$N.json; tail -4 gpurun_out/r02_lc_n$N.err | cut -c1-300
