#!/bin/bash
# 8-GPU box: lightcone strong scaling, host-link probe, standard bench at 8 ranks
mkdir -p gpurun_out
for N in 8 4 2 1; do bash scripts/gpu/r02_lc_multi.sh $N 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$"; done
for N in 1 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tools/pcie_probe_ranks.py 2>/dev/null | grep ranks
done | tee gpurun_out/r02_pcie_probe_ranks.jsonl
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_g8.json 2> gpurun_out/r02_bench_g8.err
echo "bench8 exit $?"; grep -o '"value": [0-9.]*' gpurun_out/r02_bench_g8.json | head -2
