#!/usr/bin/env bash
# round-2 first contact: 2 GPUs.  link ceiling, bench at N=1 and N=2 with the new self-explaining line
set -x
O=gpurun_out
nvidia-smi -L > $O/gpus.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  tools/pcie_ceiling.py --json $O/pcie_ceiling_n2.json > $O/pcie_n2.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus 2 --steps 3 --warmup 1 --impl reference > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err
tail -c 600 $O/bench_n1.err $O/bench_n2.err
