#!/usr/bin/env bash
# round-2, 8 GPUs: concurrent host<->device link ceiling (all subsets), then the bench line at N=8
set -x
O=gpurun_out
nvidia-smi -L > $O/gpus8.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
  tools/pcie_ceiling.py --json $O/pcie_ceiling_n8.json > $O/pcie_n8.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err
tail -c 600 $O/bench_n8.err
