#!/usr/bin/env bash
# round 2, closing 8-GPU evidence: the two-device test, the bench line at N=8, one host process driving all GPUs
set -x
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -q -k "multi or bindings" > $O/final8_pytest.log 2>&1; tail -3 $O/final8_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 8 --steps 20 --warmup 5 > $O/final8_bench_n8.json 2> $O/final8_bench_n8.err
tail -c 400 $O/final8_bench_n8.err
timeout 600 python tools/multi_e2e.py --frames 512 --json $O/final8_multi_e2e.json > $O/final8_multi_e2e.log 2>&1; cat $O/final8_multi_e2e.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus 4 --steps 20 --warmup 5 --no-also > $O/final8_bench_n4.json 2>/dev/null
