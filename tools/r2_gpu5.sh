#!/usr/bin/env bash
# round 2, fifth 1-GPU pass: ragged pooling kernel + bindings parity, finer flex tile sweep
set -x
O=gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-also"
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu5.log 2>&1; tail -15 $O/pytest_gpu5.log
timeout 900 python tools/sweep_flex.py wxga_rgb,wxga_f2,port_f1,cfg3odd,cfg4odd 256 20480,26624,28672,30720,36864,40960 2 > $O/sweep_flex_r2e.txt 2>&1; grep -v best $O/sweep_flex_r2e.txt | tail -40
