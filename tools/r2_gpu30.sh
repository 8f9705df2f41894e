#!/usr/bin/env bash
# long randomised runs of the oracle parity test (every kernel family vs the CPU oracle), three seeds
set -x
for S in 101 102 103; do
  CSIC_RANDOM_CASES=20000 CSIC_RANDOM_SEED=$S timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k randomised_parameter_space > gpurun_out/g30_random_$S.log 2>&1; tail -3 gpurun_out/g30_random_$S.log
done
