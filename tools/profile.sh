#!/usr/bin/env bash
# ncu evidence for one workload: launch list + one full capture of the row kernel (run under gpurun).
# usage: tools/profile.sh <workload> <tag>
set -u
W=$1; TAG=$2
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --workload $W --frames 64"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_${W}.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${W}_${TAG}.csv $CMD > gpurun_out/ncu_l_${W}.log 2>&1
$CMD > gpurun_out/plain2_${W}.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:csic_(rows|pool)_kernel" -s 4 -c 2 -f -o gpurun_out/prof_${W}_${TAG} $CMD > gpurun_out/ncu_f_${W}.log 2>&1
tail -n 2 gpurun_out/ncu_f_${W}.log
