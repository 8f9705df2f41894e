#!/usr/bin/env bash
# ncu full capture of the flex kernel on one workload (run under gpurun).  usage: tools/profile_flex.sh <workload> <tag> [extra bench args]
set -u
W=$1; TAG=$2; shift 2
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-verify --workload $W --frames 64 $*"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_${W}_${TAG}.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:csic_flex_kernel" -s 4 -c 1 -f -o gpurun_out/prof_${W}_${TAG} $CMD > gpurun_out/ncu_f_${W}_${TAG}.log 2>&1
tail -n 2 gpurun_out/ncu_f_${W}_${TAG}.log
